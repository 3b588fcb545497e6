#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counts with nvdisasm line info -> executed instructions per source line.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [top N]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
# locate function text
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") ):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ia, ie, isamp = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
base = int(rows[hdr + 1][ia], 16)
per = collections.Counter(); samp = collections.Counter(); tot = 0; ops = collections.Counter()
for r in rows[hdr + 1:]:
    off = int(r[ia], 16) - base
    n = int(r[ie]); tot += n
    src, txt = line_of.get(off, (None, "?"))
    per[src] += n; samp[src] += int(r[isamp] or 0)
    ops[txt.split()[0] if not txt.startswith("@") else txt.split()[1]] += n
print(f"total warp-instructions {tot}")
for (src, n) in per.most_common(top):
    print(f"{n:14d} {100.0*n/tot:5.1f}%  samples {samp[src]:7d}  {src}")
print("--- opcode mix")
for op, n in ops.most_common(25):
    print(f"{n:14d} {100.0*n/tot:5.1f}%  {op}")
