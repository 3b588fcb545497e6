#!/bin/bash
# round 2, GPU call 8: several-chunk CIGAR test, L2 prefetch A/B, approximate-max with tall panels, C3 at 100 k pairs with the lagged CIGAR drain,
# ncu capture of the warp (waves) kernel on 64 x 50 kb pairs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/r2_call8.txt
: > $O
echo "== tests" >> $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "several_chunks or fuzz_vs_oracle or ring or mid_size or c3_sample or mixed_lengths" 2>&1 | tail -4 >> $O
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d.get("parity_sample_ok"))'
run() { echo -n "$1: " >> $O; shift; env "$@" 2>&1 | python -c "$P" >> $O 2>&1; }
for l in build/ab/r6_base.so build/ab/r6_pf.so build/ab/r6_base.so build/ab/r6_pf.so; do
	run "c2 500k $l" KSW2B_LIB=$PWD/$l timeout 300 python bench.py --no-cpu --configs none --pairs 500000 --steps 3
done
for l in build/ab/r6_base.so build/ab/r6_pf.so; do
	run "c2 approx 500k $l" KSW2B_LIB=$PWD/$l timeout 300 python bench.py --approx --no-cpu --pairs 500000 --steps 3
	run "c3 20k thread $l" KSW2B_LIB=$PWD/$l KSW2B_MODE=1 timeout 600 python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1
done
echo "== approximate max, current library" >> $O
run "c2 approx 1M" timeout 600 python bench.py --approx --steps 5
run "c4 approx 1776" timeout 600 python bench.py --approx --workload c4 --pairs 1776 --steps 1
run "c3 approx 20k" timeout 600 python bench.py --approx --workload c3 --pairs 20000 --steps 1
echo "== C3 100k pairs" >> $O
run "c3 100k" timeout 900 python bench.py --workload c3 --steps 1
echo "== ncu: warp (waves) kernel, 64 x 50 kb" >> $O
KSW2B_MODE=2 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:ks_fill_warp_kernel -c 1 -o /tmp/ncu/r2_c4_warp -f python bench.py --no-cpu --workload c4 --pairs 64 --steps 1 > /tmp/ncu/c4.log 2>&1
tail -1 /tmp/ncu/c4.log | cut -c1-200 >> $O
python scripts/ncu_report.py /tmp/ncu/r2_c4_warp.ncu-rep > gpurun_out/r2_c4_warp_ncu_full.txt 2>> $O
python scripts/ncu_lines.py /tmp/ncu/r2_c4_warp.ncu-rep ksw2_b200/libksw2_b200.so ks_fill_warp_kernelILi0ELi0 45 > gpurun_out/r2_c4_warp_lines.txt 2>> $O
echo done >> $O
