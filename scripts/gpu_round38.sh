#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== ordered: parity file"; KOLIBRIE_ORDERED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== extra cfg3"; timeout 600 python bench_extra.py --only cfg3 2>&1 | grep "^{" | cut -c1-400
