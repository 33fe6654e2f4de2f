#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== pytest gpu ordered"; KOLIBRIE_ORDERED=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/b31.json; python -c "
import json; d=json.load(open('gpurun_out/b31.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']); sp=d['scan_path']; print(sp['value'], sp['ms_per_step'], {k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()}); print(d['e2e']['value'])"
