#!/usr/bin/env python
"""Per-opcode source-line breakdown of an ncu report (executed warp-instructions).
usage: ncu_ops.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> <opcode prefix> [top N]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, lib, kname, opc = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
line_of = {}; cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip()); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ia, ie = H.index("Address"), H.index("Instructions Executed")
base = int(rows[hdr + 1][ia], 16)
per = collections.Counter(); tot = 0
for r in rows[hdr + 1:]:
    off = int(r[ia], 16) - base; n = int(r[ie])
    src, txt = line_of.get(off, (None, "?"))
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    if op.startswith(opc):
        per[src] += n; tot += n
print(f"{opc}: total {tot}")
for src, n in per.most_common(top):
    print(f"{n:14d} {100.0*n/max(1,tot):5.1f}%  {src}")
