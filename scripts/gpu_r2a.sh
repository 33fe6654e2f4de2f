#!/bin/bash
# round 2, first GPU pass: new tests first (fail fast), whole suite, smoke, one-GPU bench
set -u
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_plan.py -m gpu -x -q 2>&1 | tail -15
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'host_us',d['details']['host_overhead_us_per_step'],'roof',d['roofline']['frac'],d['roofline']['ms_per_launch'])
    print('sync',d['sync_path']['value'],d['sync_path']['ms_per_step'])
    sp=d['scan_path']; print('scan',sp['value'],sp['ms_per_step'],{k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})
    print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print('10M',d.get('cfg2_10M')); print('cpu',d.get('cpu_baseline')); print(d['details']['store']); print(d['clocks'])
except Exception as e: print('parse failed',e)
PY
