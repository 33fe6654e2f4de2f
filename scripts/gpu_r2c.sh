#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== plan tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_plan.py -m gpu -x -q 2>&1 | tail -8
echo "== group check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/dist_group_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | grep "^rank\|Error\|error\|assert" | tail -12 | tee gpurun_out/dist_group_${N}gpu_r2c.txt
