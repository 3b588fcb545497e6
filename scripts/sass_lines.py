#!/usr/bin/env python
"""Static per-source-line instruction count of an address range of a kernel (e.g. a loop found by sass_loops.py).
usage: sass_lines.py <lib.so> <kernel substring> <lo hex> <hi hex> [top N]"""
import re, subprocess, os, sys, tempfile, collections
lib, kname = sys.argv[1:3]; lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16); top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
cur = None; per = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and lo <= int(m.group(1), 16) <= hi:
        t = m.group(2).split(); op = t[1] if t[0].startswith("@") else t[0]
        per[cur] += 1; ops[cur][op] += 1
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = {f: open(os.path.join(root, "ksw2_b200", "csrc", f)).read().splitlines() for f in os.listdir(os.path.join(root, "ksw2_b200", "csrc"))}
print("total", sum(per.values()))
for k, n in per.most_common(top):
    txt = src.get(k[0], [""] * 10**6)[k[1] - 1].strip()[:90] if k and k[0] in src else ""
    print(f"{n:4d} {k[0] if k else '?'}:{k[1] if k else 0:4d} {dict(ops[k].most_common(4))}  {txt}")
