#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -v -m gpu > gpurun_out/pytest_gpu.log 2>&1; grep -E "PASSED|FAILED|ERROR|Fatal|Segmentation|passed|failed" gpurun_out/pytest_gpu.log | tail -25; grep -n -B2 -A12 "Fatal Python error" gpurun_out/pytest_gpu.log | head -40
echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_full_r1c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['families']), d['e2e'])"
echo "== ncu full scan"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1c python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
