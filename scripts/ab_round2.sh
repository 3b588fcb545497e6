#!/bin/bash
# Builds the kernel variants that were prepared (and checked bit-exact with the host simulator) at the end of round 1 and A/Bs them on a
# B200.  Run the build part here (no GPU needed), then:   gpurun --timeout 1200 -- 'bash scripts/ab_round2.sh run'
#   base   = the tree as it is
#   keys   = -DKS_ARG_KEYS      arg-max position from one key max tree (interior fast step)
#   first  = -DKS_FIRST_FAST    dedicated step for the block that holds st0 (and block 0)   (measured round 2: -8 % on C2 (instruction cache), removed)
#   both   = both
set -e
cd "$(dirname "$0")/.."
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -diag-suppress 550"
if [ "$1" != "run" ]; then
	mkdir -p build/ab
	$NV -o build/ab/r2_base.so ksw2_b200/csrc/ksw2_b200.cu -ldl
	$NV -DKS_ARG_KEYS -o build/ab/r2_keys.so ksw2_b200/csrc/ksw2_b200.cu -ldl
	$NV -DKS_FIRST_FAST -o build/ab/r2_first.so ksw2_b200/csrc/ksw2_b200.cu -ldl
	$NV -DKS_ARG_KEYS -DKS_FIRST_FAST -o build/ab/r2_both.so ksw2_b200/csrc/ksw2_b200.cu -ldl
	ls -la build/ab/r2_*.so
	exit 0
fi
L="build/ab/r2_base.so build/ab/r2_keys.so build/ab/r2_first.so build/ab/r2_both.so"
echo "== C2 (500k pairs)"; REPS=2 ARGS="--no-cpu --pairs 500000 --steps 3" bash scripts/ab.sh $L
echo "== C3 (20k pairs)"; REPS=1 ARGS="--no-cpu --workload c3 --pairs 20000 --steps 2" bash scripts/ab.sh $L
echo "== parity of every variant (fuzz + golden)"
for l in $L; do echo -n "$l: "; KSW2B_LIB=$PWD/$l python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fuzz_vs_oracle or golden or c2_sample or c3_sample" 2>&1 | tail -1; done
