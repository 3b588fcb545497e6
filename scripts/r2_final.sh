#!/bin/bash
# round 2, final validation on one GPU: full gpu suite, smoke, default bench line, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_final.txt
: > $O
echo "== pytest -m gpu (all)" >> $O
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $O
echo "== smoke" >> $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $O
echo "== reference arm" >> $O
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>> $O
echo "== full default bench" >> $O
( time timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>> $O
tail -c 300 gpurun_out/r2_bench_default.err >> $O
python - >> $O 2>&1 <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print("C2", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "pageable", round(d['e2e']['pageable']['value'],1), "batch_api", round(d['e2e']['batch_api']['value'],1), "parity", d['parity_sample_ok'], "clocks", d['clocks'])
for k,v in d['configs'].items(): print(k, round(v['value'],1), "e2e", round(v['e2e']['value'],1), "parity", v['parity'], "pairs", v['pairs'], "cpu", round(v['cpu_baseline']['value'],1))
r=json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().splitlines()[-1])
print("ref", round(r['value'],1), {k:round(v['value'],1) for k,v in r['configs'].items()})
PY
echo done >> $O
