#!/bin/bash
# first GPU contact: parity tests, smoke, a short bench; everything under its own timeout
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "== pytest gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40
echo "== bench small" ; timeout 600 python bench.py --employees 1000000 --steps 5 --warmup 3 --cpu-sample 100000 2>&1 | tail -5 | tee gpurun_out/bench_small.json
