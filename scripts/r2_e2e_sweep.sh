#!/bin/bash
# e2e segmentation sweep on C2 (ksw2b_align pipeline knobs): first segment %, number of middle segments, last segment %
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_e2e_sweep.txt
: > $O
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pageable", round(d["e2e"]["pageable"]["value"],1))'
for cfg in "10 3 6" "5 3 6" "5 5 4" "8 4 4" "5 6 3" "3 6 3" "10 3 6"; do set -- $cfg
	echo -n "first=$1% rest=$2 last=$3%: " >> $O
	KSW2B_FIRST_PCT=$1 KSW2B_REST_SEGS=$2 KSW2B_LAST_PCT=$3 timeout 200 python bench.py --no-cpu --configs none --steps 3 2>&1 | python -c "$P" >> $O 2>&1
done
echo done >> $O
