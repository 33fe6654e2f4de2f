#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_fast -s 1 -c 1 -o gpurun_out/prof_probe_index_r1j -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/prof_probe_index_r1j.ncu-rep
