#!/bin/bash
# round 2, GPU call 9: L2 persisting window A/B on the 150 bp workload (exact and approximate)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call9.txt
: > $O
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity_sample_ok"))'
run() { echo -n "$1: " >> $O; shift; env "$@" 2>&1 | python -c "$P" >> $O 2>&1; }
for v in 0 1 0 1; do run "c2 1M l2persist=$v" KSW2B_L2PERSIST=$v timeout 300 python bench.py --no-cpu --configs none --steps 5; done
for v in 0 1; do run "c2 approx 1M l2persist=$v" KSW2B_L2PERSIST=$v timeout 300 python bench.py --approx --no-cpu --steps 5; done
for v in 0 1; do run "c3 20k l2persist=$v" KSW2B_L2PERSIST=$v timeout 300 python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1; done
echo "== ncu dram bytes with the window" >> $O
KSW2B_L2PERSIST=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ks_fill_kernel -s 3 -c 1 python bench.py --no-cpu --configs none --pairs 300000 --steps 1 --warmup 3 2>&1 | grep -E "dram__|gpu__time" >> $O
echo done >> $O
