#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:derive_kernel -s 10 -c 2 -o gpurun_out/prof_derive_r1q -f python scripts/datalog_scale.py 48888890 > gpurun_out/ncu_derive.log 2>&1
ls -la gpurun_out/prof_derive_r1q.ncu-rep; tail -2 gpurun_out/ncu_derive.log | cut -c1-300
