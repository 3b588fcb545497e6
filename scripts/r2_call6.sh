#!/bin/bash
# round 2, GPU call 6a: full gpu suite, smoke, default bench line + reference arm, C2 tuning sweep, combining layer (text results only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call6.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O 2>&1
echo "== pytest -m gpu (all)" >> $O
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $O
echo "== smoke" >> $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $O
echo "== reference arm" >> $O
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>> $O
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 400 gpurun_out/r2_bench_default.err >> $O
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1))'
echo "== C2 tuning sweep (panel threads ctas)" >> $O
for cfg in "15 96 4" "18 96 4" "12 96 4" "18 128 3" "14 128 3" "20 64 6" "15 96 4"; do set -- $cfg; echo -n "panel=$1 threads=$2 ctas=$3: " >> $O; timeout 300 python bench.py --no-cpu --configs none --pairs 500000 --steps 3 --panel $1 --threads $2 --ctas $3 2>&1 | python -c "$P" >> $O 2>&1; done
echo "== exts2 throughput" >> $O
timeout 600 python scripts/exts2_run.py 2>&1 | tail -1 >> $O
echo "== combining layer" >> $O
timeout 600 python scripts/combine_bench.py 100000 1,16,64,256,1024 >> $O 2>&1
echo done >> $O
