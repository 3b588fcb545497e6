#!/usr/bin/env python
"""Aggregate scripts/ncu_lines.py output (stdin) into code regions of ksw2_tile.cuh, per block-step.
usage: ncu_lines.py ... 400 | ncu_regions.py <block-steps (warp level)>"""
import sys, re, collections
steps = float(sys.argv[1])
src = open(__file__.replace("scripts/ncu_regions.py", "ksw2_b200/csrc/ksw2_tile.cuh")).read().splitlines()
marks = [("ks_imax", "helpers: geo/imax/bnd/zdrop"), ("struct KsBlk", "blk/pack/mask helpers"), ("void ks_score_row", "score row+qshift+qload"), ("void ks_splice", "splice"),
         ("define KS_NOCAND", "hmax/harg/scratch helpers"), ("struct KsTile", "tile_begin"), ("KS_HD bool ks_tile_step", "step prologue"), ("---- core: all 16 lanes", "core"),
         ("---- exact max: H[]", "H: top pre + update + insert"), ("// block maximum over the SIMD-part", "block max/arg/merge/hst0"), ("if (!is_top) bout", "finalize+cout"),
         ("KS_HD void ks_tile_end", "tile_end"), ("KS_HD void ks_tile(", "ks_tile loop")]
bounds = []
for key, name in marks:
    ln = next(i + 1 for i, l in enumerate(src) if key in l)
    bounds.append((ln, name))
bounds.sort()
tot = collections.Counter(); other = collections.Counter(); T = 0
for l in sys.stdin:
    m = re.match(r"\s*(\d+)\s+[\d.]+%\s+samples\s+(\d+)\s+\('(.*)', (\d+)\)", l)
    if not m:
        continue
    n = int(m.group(1)); f = m.group(3); ln = int(m.group(4)); T += n
    if f == "ksw2_tile.cuh":
        name = "top-of-file"
        for b, nm in bounds:
            if ln >= b:
                name = nm
        tot[name] += n
    else:
        other[f] += n
for b, nm in bounds:
    print(f"{tot[nm]/steps:7.1f}  {nm}")
for f, n in other.most_common():
    print(f"{n/steps:7.1f}  {f}")
print(f"{T/steps:7.1f}  total")
