#!/bin/bash
# refreshes everything profiles/ is generated from, in two calls (a call may bring back at most 64 MiB; a full capture is ~19 MiB):
#   gpurun -- 'bash scripts/gpu_profiles.sh TAG a'   launch list + index probe + star scan
#   gpurun -- 'bash scripts/gpu_profiles.sh TAG b'   scan-path probe + the two passes of the Datalog candidate dedup
# summaries: python scripts/summarize_ncu.py TAG
set -u
mkdir -p gpurun_out
TAG=${1:-r2z}
PART=${2:-a}
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-adversarial"
if [ "$PART" = a ]; then
  echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_bench.log 2>&1
  # index path = probe_index_kernel (table mode); scan path = scan_star_kernel + probe_fast_kernel
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_index -s 3 -c 1 -o gpurun_out/prof_probe_index_$TAG -f $B > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_star -s 1 -c 1 -o gpurun_out/prof_scan_$TAG -f $B > /dev/null 2>&1
else
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_fast -s 1 -c 1 -o gpurun_out/prof_probe_$TAG -f $B > /dev/null 2>&1
  # Datalog candidate dedup at 1/5 of config 4 (ncu replays every pass ~40 times, saving and restoring the table each time)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:derive_probe -s 3 -c 1 -o gpurun_out/prof_derive_$TAG -f python scripts/datalog_trace.py 10000000 > /dev/null 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:derive_partition -s 3 -c 1 -o gpurun_out/prof_derivepart_$TAG -f python scripts/datalog_trace.py 10000000 > /dev/null 2>&1
fi
ls -la gpurun_out | grep $TAG
