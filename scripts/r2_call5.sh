#!/bin/bash
# round 2, GPU call 5: ring schedule / CTA mode tests and A/B by mode on the banded CIGAR workloads; C1 on the CTA kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call5.txt
: > $O
echo "== new gpu tests" >> $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or cta or warp_mode or golden or fuzz_vs_oracle or c3_sample or mixed_lengths or c5_geometry" 2>&1 | tail -8 >> $O
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity_sample_ok"))'
run() { echo -n "$1: " >> $O; shift; env "$@" 2>&1 | python -c "$P" >> $O 2>&1; }
echo "== C3 20k pairs by mode (1 thread, 2 warp/waves, 4 ring, 0 auto)" >> $O
for m in 1 2 4 0; do run "mode=$m" KSW2B_MODE=$m timeout 600 python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1; done
echo "== C3 100k pairs (auto)" >> $O
run "auto" timeout 900 python bench.py --workload c3 --steps 1
run "thread" KSW2B_MODE=1 timeout 900 python bench.py --no-cpu --workload c3 --steps 1
echo "== C5 200k pairs" >> $O
run "auto" timeout 900 python bench.py --workload c5 --steps 1
run "thread" KSW2B_MODE=1 timeout 900 python bench.py --no-cpu --workload c5 --steps 1
echo "== C1" >> $O
run "auto (cta)" timeout 300 python bench.py --workload c1
run "warp" KSW2B_MODE=2 timeout 300 python bench.py --no-cpu --workload c1
echo "== C2 / C4 sanity" >> $O
run "c2 500k" timeout 300 python bench.py --no-cpu --configs none --pairs 500000 --steps 3
run "c4 1776" timeout 300 python bench.py --no-cpu --workload c4 --pairs 1776 --steps 1
echo "== batch api e2e (C2 default headline only)" >> $O
timeout 600 python bench.py --configs none 2>&1 | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), json.dumps(d["e2e"]))' >> $O 2>&1
echo "== ncu ring kernel (C3, 4000 pairs)" >> $O
KSW2B_MODE=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_fill_ring_kernel -c 1 -o gpurun_out/r2_c3_ring -f python bench.py --no-cpu --workload c3 --pairs 4000 --steps 1 > gpurun_out/ncu_c3_ring.log 2>&1
tail -2 gpurun_out/ncu_c3_ring.log >> $O
echo done >> $O
