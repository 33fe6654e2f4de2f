#!/bin/bash
# refreshes everything profiles/ is generated from (run with: gpurun -- 'bash scripts/gpu_profiles.sh TAG')
set -u
mkdir -p gpurun_out
TAG=${1:-r1z}
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_full_$TAG.json; python -c "
import json; d=json.load(open('gpurun_out/bench_full_$TAG.json')); print(d['value'], d['ms_per_step'], d['scan_path']['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_$TAG.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
# index path = probe_index_kernel; scan path = scan_kernel + probe_fast_kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_index -s 1 -c 1 -o gpurun_out/prof_probe_index_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/prof_scan_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_fast -s 1 -c 1 -o gpurun_out/prof_probe_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
echo "== extra"; timeout 1500 python bench_extra.py --cpu 2>&1 | grep "^{" > gpurun_out/bench_extra_$TAG.jsonl; cut -c1-160 gpurun_out/bench_extra_$TAG.jsonl
ls gpurun_out | grep $TAG
