#!/usr/bin/env python
"""Per-step cost of the fill kernel from one ncu capture: executed warp-instructions inside the fast-step loop and inside the
generic step (address ranges from scripts/sass_loops.py), per source line and per opcode.
usage: ncu_steps.py <src.csv from `ncu --page source --csv`> <lib.so> <kernel substring> <fast_lo> <fast_hi> <gen_lo> <gen_hi> [lines|ops]"""
import csv, re, subprocess, os, sys, tempfile, collections
srcf, lib, kname = sys.argv[1:4]
flo, fhi, glo, ghi = [int(x, 16) for x in sys.argv[4:8]]
mode = sys.argv[8] if len(sys.argv) > 8 else "lines"
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
line_of, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srcf)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ia, ie, isrc = H.index("Address"), H.index("Instructions Executed"), H.index("Source")
base = int(rows[hdr + 1][ia], 16)
data = [(int(r[ia], 16) - base, int(r[ie]), r[isrc].strip()) for r in rows[hdr + 1:]]
tot = sum(n for _, n, _ in data)
fast_it = next(n for o, n, _ in data if o == flo)
vi = sum(n for o, n, s in data if s.split()[0] == "VIADD.16x2" or (s.startswith("@") and s.split()[1] == "VIADD.16x2"))
nfast_vi = sum(n for o, n, s in data if flo <= o <= fhi and "VIADD.16x2" in s)
vi_per_step = nfast_vi / fast_it
steps = vi / vi_per_step
gen_it = steps - fast_it
fast_n = sum(n for o, n, _ in data if flo <= o <= fhi)
gen_n = sum(n for o, n, _ in data if glo <= o <= ghi and not (flo <= o <= fhi))
print(f"total {tot}; block-steps {steps:.0f} (fast {fast_it}, generic {gen_it:.0f}); avg {tot/steps:.1f}/step")
print(f"fast step: {fast_n/fast_it:.1f} instr; generic step: {gen_n/gen_it:.1f} instr; outside the step loops: {(tot-fast_n-gen_n)/steps:.1f} per step")
def opof(s):
    t = s.split(); return t[1] if t[0].startswith("@") else t[0]
for name, lo, hi, excl, it in (("fast", flo, fhi, None, fast_it), ("generic", glo, ghi, (flo, fhi), gen_it)):
    per = collections.Counter()
    for o, n, s in data:
        if lo <= o <= hi and not (excl and excl[0] <= o <= excl[1]):
            per[line_of.get(o) if mode == "lines" else opof(s)] += n
    print(f"--- {name} step, per {mode} (instr per step)")
    for k, v in per.most_common(28): print(f"{v/it:7.1f}  {k}")
