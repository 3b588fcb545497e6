#!/usr/bin/env python
"""Static view of a kernel's SASS: every loop (backward branch) with its instruction count and opcode mix.
usage: sass_loops.py <lib.so> <mangled-name substring>"""
import re, subprocess, sys
lib, key = sys.argv[1:3]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
ins, on = [], False
for l in txt:
    if "Function :" in l:
        on = key in l
        if on: print(l.strip())
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("total", len(ins))
def opof(x):
    t = x.split()
    return t[1] if t[0].startswith('@') else t[0]
for i, (a, t) in enumerate(ins):
    if 'BRA' not in t: continue
    m = re.search(r'0x([0-9a-f]+)', t)
    if not m: continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr:
        body = ins[addr[tgt]:i + 1]
        ops = {}
        for _, x in body: ops[opof(x)] = ops.get(opof(x), 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:10]
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instrs", top)
