#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 1 0; do
  echo "== bench KOLIBRIE_PROBE_TABLE=$v"; KOLIBRIE_PROBE_TABLE=$v timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/b_tab$v.json 2> gpurun_out/b_tab$v.err; tail -c 400 gpurun_out/b_tab$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_tab$v.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'roof',d['roofline']['frac'],d['roofline']['ms_per_launch'])
a=d.get('adversarial'); print('adversarial',a and {k:a[k] for k in ('index_path','scan_path')})
PY
done
