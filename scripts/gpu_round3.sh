#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== debug group"; timeout 120 python scripts/debug_group.py 2>&1 | tail -12
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_full_r1b.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['families']), d['e2e'])"
echo "== ncu full scan"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1b python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
