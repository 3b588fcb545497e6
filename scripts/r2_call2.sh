#!/bin/bash
# round 2, GPU call 2: new tests, unroll A/B, the full default bench line (all five configs), ncu capture of the C2 fill kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call2.txt
: > $O
echo "== new gpu tests" >> $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "band_per_pair or c5_geometry or golden_thread_mode or multi_device or one_live_plan or kalloc or invalid" 2>&1 | tail -15 >> $O
L="build/ab/r2b_base.so build/ab/r2b_u2.so build/ab/r2b_u4.so"
echo "== C2 (500k pairs)" >> $O; REPS=2 ARGS="--no-cpu --configs none --pairs 500000 --steps 3" bash scripts/ab.sh $L >> $O 2>&1
echo "== C3 (20k pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c3 --pairs 20000 --steps 2" bash scripts/ab.sh $L >> $O 2>&1
echo "== C4 (592 pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c4 --pairs 592 --steps 1" bash scripts/ab.sh $L >> $O 2>&1
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 600 gpurun_out/r2_bench_default.err >> $O
echo "== reference arm" >> $O
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>> $O
echo "== ncu C2" >> $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_fill_kernel -s 3 -c 1 -o gpurun_out/r2_c2_base -f python bench.py --no-cpu --configs none --pairs 300000 --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log >> $O
echo done >> $O
