#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r1i
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_pairs_filtered -s 8 -c 2 -o gpurun_out/prof_build_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
grep -c . gpurun_out/launches_$TAG.csv
