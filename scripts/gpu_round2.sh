#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== debug group"; timeout 120 python scripts/debug_group.py 64 2>&1 | tail -12
echo "== sanitizer group"; timeout 300 compute-sanitizer --tool memcheck python scripts/debug_group.py 64 2>&1 | grep -v "^=========     at\|^=========     by" | tail -30
echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_full_r1a.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full scan"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1a python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
echo "== ncu full probe"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_direct -s 3 -c 1 -o gpurun_out/prof_probe_r1a python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
echo "== ncu full build"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_direct -s 6 -c 1 -o gpurun_out/prof_build_r1a python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out
