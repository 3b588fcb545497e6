#!/bin/bash
# round 2, GPU call 1: state of the tree on today's box + A/B of the variants prepared at the end of round 1 + occupancy variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call1.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O 2>&1
echo "== pytest -m gpu" >> $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $O
L="build/ab/r2_base.so build/ab/r2_keys.so build/ab/r2_first.so build/ab/r2_both.so"
echo "== C2 (500k pairs)" >> $O; REPS=2 ARGS="--no-cpu --pairs 500000 --steps 3" bash scripts/ab.sh $L >> $O 2>&1
echo "== C3 (20k pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c3 --pairs 20000 --steps 2" bash scripts/ab.sh $L >> $O 2>&1
echo "== occupancy variants, C2" >> $O
ARGS="--no-cpu --pairs 500000 --steps 3" bash scripts/ab2.sh "build/ab/r2_base.so 15 96 4" "build/ab/r2_lb128x4.so 13 128 4" "build/ab/r2_lb96x5.so 14 96 5" "build/ab/r2_lb96x5.so 11 96 5" "build/ab/r2_lb128x4.so 10 128 4" "build/ab/r2_base.so 15 96 4" >> $O 2>&1
echo "== C4 (592 pairs)" >> $O
KSW2B_LIB=$PWD/build/ab/r2_base.so timeout 600 python bench.py --no-cpu --workload c4 --steps 1 --warmup 3 2>&1 | tail -1 | cut -c1-400 >> $O
echo "== parity of every variant (fuzz + golden)" >> $O
for l in build/ab/r2_keys.so build/ab/r2_first.so build/ab/r2_both.so build/ab/r2_lb128x4.so; do echo -n "$l: " >> $O; KSW2B_LIB=$PWD/$l timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fuzz_vs_oracle or golden or c2_sample or c3_sample" 2>&1 | tail -1 >> $O; done
echo done >> $O
