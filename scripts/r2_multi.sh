#!/bin/bash
# round 2, multi-GPU run: usage r2_multi.sh N   (under gpurun --gpus N)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2_multi_n$N.txt
: > $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader >> $O 2>&1
echo "== device set from one caller (pytest, C example)" >> $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_device" 2>&1 | tail -2 >> $O
gcc -std=c99 -Iinclude examples/multi_gpu.c -Lksw2_b200 -lksw2_b200 -Wl,-rpath,$PWD/ksw2_b200 -o /tmp/multi_gpu && timeout 300 /tmp/multi_gpu $N >> $O 2>&1
echo "== bench, $N ranks" >> $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>> $O
tail -c 500 gpurun_out/r2_bench_n$N.err >> $O
python - >> $O 2>&1 <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print("C2", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "parity", d['parity_sample_ok'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), "e2e", round(v['e2e']['value'],1), "parity", v['parity'], "pairs", v['pairs'])
print("c_api_multi", json.dumps(d.get('c_api_multi'))[:900])
PY
echo done >> $O
