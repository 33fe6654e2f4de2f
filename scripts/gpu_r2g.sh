#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== datalog tests"; timeout 900 python -m pytest tests/test_gpu_datalog.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
echo "== cfg4 trace (partitioned)"; KOLIBRIE_TRACE=1 timeout 600 python scripts/datalog_trace.py 2> gpurun_out/trace_part.txt | tail -3; awk '/==== run 2/{f=1} f' gpurun_out/trace_part.txt | grep -v "views" | tail -45
