#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== default bench"; SECONDS=0; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -3 gpurun_out/bench_r2i.err | cut -c1-300; echo "wall ${SECONDS}s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2i.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'roof',d['roofline']['frac'])
    for k,v in d.get('other_configs',{}).items(): print(k, {a:b for a,b in v.items() if a not in ('workload','parity','roofline')}, v['roofline'].get('frac'))
    print('adv', d['adversarial']['index_path']['frac_of_peak']); print('10M',d['cfg2_10M']['value'], d['cfg2_10M']['e2e']['value']); print('cpu',d['cpu_baseline']['value'])
except Exception as e: print('parse failed',e)
PY
echo "== reference arm"; SECONDS=0; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -2 | cut -c1-900; echo "wall ${SECONDS}s"
