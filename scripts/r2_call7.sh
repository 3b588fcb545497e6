#!/bin/bash
# round 2, GPU call 6b: ncu launch list and --set full captures of the fill kernels, summarised ON THE BOX (the reports are too large to bring back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/r2_call7.txt
: > $O
echo "== ncu launch list" >> $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-cpu --configs none --steps 2 --warmup 1 > /tmp/ncu/launches.log 2>&1
tail -1 /tmp/ncu/launches.log | cut -c1-300 >> $O
cap() {   # name kernel-regex lib-kernel-substr pairs command...
	local name=$1 rx=$2 sub=$3; shift 3
	timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o /tmp/ncu/$name -f "$@" > /tmp/ncu/$name.log 2>&1
	tail -1 /tmp/ncu/$name.log | cut -c1-200 >> $O
	python scripts/ncu_report.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> $O
	python scripts/ncu_lines.py /tmp/ncu/$name.ncu-rep ksw2_b200/libksw2_b200.so $sub 45 > gpurun_out/${name}_lines.txt 2>> $O
}
cap r2_c2_thread "ks_fill_kernel" "ks_fill_kernelILi0ELi0" env KSW2B_X=1 python bench.py --no-cpu --configs none --pairs 300000 --steps 1 --warmup 0
cap r2_c3_thread "ks_fill_kernel" "ks_fill_kernelILi1ELi1" env KSW2B_MODE=1 python bench.py --no-cpu --workload c3 --pairs 20000 --steps 1
cap r2_c3_ring "ks_fill_ring_kernel" "ks_fill_ring_kernelILi1ELi1" env KSW2B_MODE=4 python bench.py --no-cpu --workload c3 --pairs 4000 --steps 1
cap r2_c4_warp "ks_fill_warp_kernel" "ks_fill_warp_kernelILi0ELi0" python bench.py --no-cpu --workload c4 --pairs 64 --steps 1
cap r2_exts2_thread "ks_fill_kernel" "ks_fill_kernelILi2ELi0" python scripts/exts2_run.py
cap r2_c1_cta "ks_fill_cta_kernel" "ks_fill_cta_kernelILi0ELi1" python bench.py --no-cpu --workload c1
cap r2_c2_approx "ks_fill_kernel" "ks_fill_kernelILi0ELi4" python bench.py --approx --no-cpu --pairs 300000 --steps 1 --warmup 0
python scripts/ncu_traffic.py c2=/tmp/ncu/r2_c2_thread.ncu-rep:300000:profiles/r2_c2_thread_ncu_full.txt c3=/tmp/ncu/r2_c3_thread.ncu-rep:20000:profiles/r2_c3_thread_ncu_full.txt \
	c4=/tmp/ncu/r2_c4_warp.ncu-rep:64:profiles/r2_c4_warp_ncu_full.txt c1=/tmp/ncu/r2_c1_cta.ncu-rep:1:profiles/r2_c1_cta_ncu_full.txt > gpurun_out/r2_traffic.json 2>> $O
cp /tmp/ncu/r2_c2_thread.ncu-rep gpurun_out/ 2>/dev/null
ls -la /tmp/ncu gpurun_out >> $O
echo done >> $O
