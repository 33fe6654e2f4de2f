#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== pytest gpu ordered"; KOLIBRIE_ORDERED=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== rsp"; KOLIBRIE_TRACE=1 timeout 300 python scripts/rsp_breakdown.py 2>&1 | tail -5
echo "== bench (e2e uses chunked segments)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/b29.json; python -c "
import json; d=json.load(open('gpurun_out/b29.json')); print(d['value'], d['ms_per_step'], d['scan_path']['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
