#!/bin/bash
# round 2, GPU call 10: tail split test, C5 at one GPU with a rank's share of 8 (strong-scaling regime), C5 full, default bench (pinned result buffer)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call10.txt
: > $O
echo "== tests" >> $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_between or several_chunks or mixed_lengths or c5_geometry or multi_device" 2>&1 | tail -4 >> $O
echo "== C5: the share of rank 0 of 8 ranks, on one GPU (tail split on / forced thread mode)" >> $O
python - >> $O 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, bench, harness as H, ksw2_b200 as K
mat = H.simple_mat(5, 2, 4)
bs = bench.build_batches("c5", 0, 8)
for mode in (0, 1, 0):
    ctx = K.Context(0); ctx.set_mode(mode, 0); ctx.set_timing(True)
    tot = 0.0; cells = 0
    for rep in range(2):
        tot = 0.0; cells = 0
        for b in bs:
            r, _ = ctx.align_packed(K.make_params(b.kind, mat, **b.par), b.qcat, b.qoff, b.tcat, b.toff, None, b.w, want_cigars=False)
            f, nf, sp, nl = ctx.last_timing(); tot += sp
            cells += int(bench.cells_lanes(b.qoff, b.toff, b.w, r["n_diag"])[0].sum())
    print(f"mode {mode}: {cells / tot / 1e6:.1f} GCUPS, device span {tot:.1f} ms for {sum(b.n for b in bs)} pairs")
    ctx.close()
PY
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 300 gpurun_out/r2_bench_default.err >> $O
python - >> $O 2>&1 <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print("C2", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "pageable", round(d['e2e']['pageable']['value'],1), "batch_api", round(d['e2e']['batch_api']['value'],1), "parity", d['parity_sample_ok'], "traffic", d['roofline']['traffic'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), "e2e", round(v['e2e']['value'],1), "parity", v['parity'], "pairs", v['pairs'], "cpu", round(v['cpu_baseline']['value'],1))
PY
echo done >> $O
