#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench_extra 1 GPU"; timeout 1500 python bench_extra.py --cpu 2>&1 | grep "^{" | tee gpurun_out/bench_extra_1gpu.jsonl | cut -c1-400
echo "== bench_extra 2 GPU cfg3,cfg4"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench_extra.py --only cfg3,cfg4 2>&1 | grep "^{" | tee gpurun_out/bench_extra_2gpu.jsonl | cut -c1-400
