#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== pytest gpu, chained joins"; KOLIBRIE_CSR_JOIN=0 timeout 1200 python -m pytest tests -m gpu -x -q -k "join or datalog or fc or taxonomy or rsp or bgp" 2>&1 | tail -3
echo "== pytest gpu ordered"; KOLIBRIE_ORDERED=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
echo "== cfg4 CSR_JOIN=$v"; KOLIBRIE_CSR_JOIN=$v timeout 900 python scripts/datalog_scale.py 48888890 2>&1 | tail -2 | cut -c1-1200
done
