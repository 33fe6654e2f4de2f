#!/bin/bash
set -u
mkdir -p gpurun_out
for N in 1000000 10000000; do echo "== datalog $N"; timeout 900 python scripts/datalog_scale.py $N 2>&1 | tail -2 | tee -a gpurun_out/datalog_scale.jsonl; done
echo "== datalog 48.9M"; timeout 1200 python scripts/datalog_scale.py 48888890 2>&1 | tail -2 | tee -a gpurun_out/datalog_scale.jsonl
