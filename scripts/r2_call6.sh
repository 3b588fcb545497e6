#!/bin/bash
# round 2, GPU call 6: the evidence run -- full gpu suite, smoke, default bench line + reference arm, C2 tuning sweep, ncu launch list and
# --set full captures of the four fill kernels, combining layer
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
echo "== combining layer" >> $O
timeout 600 python scripts/combine_bench.py 100000 1,16,64,256,1024 >> $O 2>&1
echo "== ncu launch list" >> $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-cpu --configs none --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200 >> $O
echo "== ncu --set full: C2 thread kernel, C3 thread kernel, C4 warp kernel (64 pairs), exts2 (2000 x 1 kb), C1 cta kernel" >> $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_fill_kernel -s 3 -c 1 -o gpurun_out/r2_c2_final -f python bench.py --no-cpu --configs none --pairs 300000 --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log >> $O
KSW2B_MODE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_fill_kernel -c 1 -o gpurun_out/r2_c3_thread -f python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1 > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_fill_warp_kernel -c 1 -o gpurun_out/r2_c4_warp -f python bench.py --no-cpu --workload c4 --pairs 64 --steps 1 > gpurun_out/ncu_c4.log 2>&1; tail -1 gpurun_out/ncu_c4.log >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_fill_kernel -c 1 -o gpurun_out/r2_exts2 -f python scripts/exts2_run.py > gpurun_out/ncu_exts2.log 2>&1; tail -2 gpurun_out/ncu_exts2.log >> $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_fill_cta_kernel -c 1 -o gpurun_out/r2_c1_cta -f python bench.py --no-cpu --workload c1 > gpurun_out/ncu_c1.log 2>&1; tail -1 gpurun_out/ncu_c1.log >> $O
echo done >> $O
