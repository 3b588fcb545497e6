#!/bin/bash
# round 2, GPU call 4: full gpu suite; old (r3_base) vs new (approximate-max kernel variants, converged mixed warp step, per-chunk mode, per-kind
# launch bounds, coded-target reload for short pairs) on every configuration; approximate-max throughput; default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call4.txt
: > $O
echo "== pytest -m gpu (all)" >> $O
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 >> $O
L="build/ab/r3_base.so build/ab/r4_new.so"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity_sample_ok"))'
echo "== C2 (500k pairs)" >> $O; REPS=2 ARGS="--no-cpu --configs none --pairs 500000 --steps 3" bash scripts/ab.sh $L >> $O 2>&1
echo "== C3 (20k pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c3 --pairs 20000 --steps 2" bash scripts/ab.sh $L >> $O 2>&1
echo "== C3 (20k pairs) warp mode forced" >> $O; for l in $L; do echo -n "$l: " >> $O; KSW2B_MODE=2 KSW2B_LIB=$PWD/$l timeout 600 python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1 2>&1 | python -c "$P" >> $O 2>&1; done
echo "== C4 (1776 pairs)" >> $O; REPS=1 ARGS="--no-cpu --workload c4 --pairs 1776 --steps 1" bash scripts/ab.sh $L >> $O 2>&1
echo "== C5 (200k pairs)" >> $O; for l in $L; do echo -n "$l: " >> $O; KSW2B_LIB=$PWD/$l timeout 900 python bench.py --workload c5 --steps 1 2>&1 | python -c "$P" >> $O 2>&1; done
echo "== C1" >> $O; for l in $L; do echo -n "$l: " >> $O; KSW2B_LIB=$PWD/$l timeout 300 python bench.py --workload c1 2>&1 | python -c "$P" >> $O 2>&1; done
echo "== approximate max (KSW_EZ_APPROX_MAX), new library; then the scalar kernel of round 1 on a small sample" >> $O
echo -n "c2 approx: " >> $O; timeout 600 python bench.py --approx --pairs 500000 --steps 3 2>&1 | python -c "$P" >> $O 2>&1
echo -n "c4 approx (1776 pairs): " >> $O; timeout 600 python bench.py --approx --workload c4 --pairs 1776 --steps 1 2>&1 | python -c "$P" >> $O 2>&1
echo -n "c3 approx (20k pairs): " >> $O; timeout 600 python bench.py --approx --workload c3 --pairs 20000 --steps 1 2>&1 | python -c "$P" >> $O 2>&1
echo -n "c2 approx, scalar kernel (KSW2B_SCALAR_APPROX=1, 100k pairs): " >> $O; KSW2B_SCALAR_APPROX=1 timeout 600 python bench.py --approx --no-cpu --pairs 100000 --steps 1 2>&1 | python -c "$P" >> $O 2>&1
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 600 gpurun_out/r2_bench_default.err >> $O
echo done >> $O
