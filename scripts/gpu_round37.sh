#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/dist_shuffle_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL INFO\|warn" | tail -25
