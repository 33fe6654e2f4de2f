#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== extra"; timeout 900 python bench_extra.py --only cfg1,cfg3 2>&1 | grep "^{" | cut -c1-420
