#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (star scan kernel on)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_rsp.py tests/test_gpu_plan.py -m gpu -x -q 2>&1 | tail -4
for v in 1 0; do
  echo "== bench KOLIBRIE_SCAN_STAR=$v"; KOLIBRIE_SCAN_STAR=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_star$v.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_star$v.json').read())
sp=d['scan_path']; print('scan',sp['value'],sp['ms_per_step'],{k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})
PY
done
echo "== inst count"; timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:scan_star -s 2 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -E "scan_star|inst_executed|time_duration|issue_active" | head -8
