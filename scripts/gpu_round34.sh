#!/bin/bash
set -u
mkdir -p gpurun_out
for g in 0 32 64 128; do
echo "== L2 fetch $g"
KOLIBRIE_L2_FETCH=$g timeout 600 python scripts/datalog_trace.py 2>&1 | grep "wall\|L2 fetch" | tail -4
KOLIBRIE_L2_FETCH=$g timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/l2_$g.json; python -c "
import json; d=json.load(open('gpurun_out/l2_$g.json')); sp=d['scan_path']; print(d['value'], d['roofline']['ms_per_launch'], {k:round(v['ms'],4) for k,v in sp['roofline']['families'].items()})"
done
