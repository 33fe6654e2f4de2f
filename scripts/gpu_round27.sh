#!/bin/bash
# 8-GPU run of the bench contract exactly as the driver launches it
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== bench N=$N"; timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${N}gpu_r1p.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_r1p.json')); print(d['value'], d['ms_per_step'], d.get('scan_path',{}).get('value'), d['e2e']['value'], d['n_gpus'], d['clocks'])"
echo "== ref arm N=$N"; timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-200
