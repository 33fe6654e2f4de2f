#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (default=unordered)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== bench 2 gpus"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_2gpu_r1.json | cut -c1-1500
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_r1.json | cut -c1-600
