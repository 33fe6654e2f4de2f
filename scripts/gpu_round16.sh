#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== pytest gpu ordered"; KOLIBRIE_ORDERED=1 timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_full_r1h.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('INDEX', d['value'], d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'] or 0,3)) for k,v in d['roofline']['families'].items()}, d['roofline']['device_ms_per_step'], d['config']['store'][:80])
sp=d['scan_path']; print('SCAN', sp['value'], sp['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})
print('E2E', d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'])"
