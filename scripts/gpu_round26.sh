#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/datalog_trace.py 2>&1 | tail -6
echo "== extra"; timeout 900 python bench_extra.py --cpu 2>&1 | grep "^{" > gpurun_out/bench_extra_r1p.jsonl; cut -c1-330 gpurun_out/bench_extra_r1p.jsonl
