#!/bin/bash
# round 2, GPU call 3: full gpu suite on the new library, prefetch / state-size A/B, warp-mode panel sweep on C4, default bench line, combining layer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call3.txt
: > $O
echo "== pytest -m gpu (all)" >> $O
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $O
L="build/ab/r3_base.so build/ab/r3_qp2.so build/ab/r3_tr.so build/ab/r3_both.so"
echo "== C2 (500k pairs)" >> $O; REPS=2 ARGS="--no-cpu --configs none --pairs 500000 --steps 3" bash scripts/ab.sh $L >> $O 2>&1
echo "== C3 (20k pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c3 --pairs 20000 --steps 2" bash scripts/ab.sh $L >> $O 2>&1
echo "== C4 (592 pairs), warp panel sweep" >> $O
for wp in 128 512 1024 4096; do echo -n "wpanel=$wp: " >> $O; KSW2B_WPANEL=$wp KSW2B_LIB=$PWD/build/ab/r3_base.so timeout 300 python bench.py --no-cpu --workload c4 --pairs 592 --steps 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" >> $O 2>&1; done
echo -n "C4 1776 pairs wpanel=1024: " >> $O; KSW2B_LIB=$PWD/build/ab/r3_base.so timeout 300 python bench.py --no-cpu --workload c4 --pairs 1776 --steps 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" >> $O 2>&1
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 600 gpurun_out/r2_bench_default.err >> $O
echo "== combining layer" >> $O
timeout 600 python scripts/combine_bench.py 100000 1,16,64,256,1024 >> $O 2>&1
KSW2B_LANES=1 timeout 300 python scripts/combine_bench.py 100000 64,256 >> $O 2>&1
KSW2B_LINGER_US=50 timeout 300 python scripts/combine_bench.py 100000 64,256,1024 >> $O 2>&1
echo done >> $O
