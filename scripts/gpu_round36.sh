#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu datalog"; timeout 1200 python -m pytest tests -m gpu -x -q -k "datalog or fc or taxonomy or rsp or cpp or known" 2>&1 | tail -3
for v in kolibrie_b200 kb_u1; do
echo "== $v"; KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/lib$v.so KOLIBRIE_TRACE=1 timeout 600 python scripts/datalog_trace.py 2>&1 | tail -150 > gpurun_out/dl_trace_$v.txt; grep "wall" gpurun_out/dl_trace_$v.txt
done
