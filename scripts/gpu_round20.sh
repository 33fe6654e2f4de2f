#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== ordered index tests"; KOLIBRIE_ORDERED=1 timeout 600 python -m pytest tests -m gpu -x -q -k "index" 2>&1 | tail -3
for v in 0 1; do
echo "== bench INDEX_KERNEL=$v"; KOLIBRIE_INDEX_KERNEL=$v timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/ab_$v.json; python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
