#!/bin/bash
# the bench contract under torchrun exactly as the driver launches it: bash scripts/gpu_multi.sh N TAG
set -u
mkdir -p gpurun_out
N=${1:-8}; TAG=${2:-r1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${N}gpu_$TAG.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_$TAG.json')); print(d['value'], d['ms_per_step'], d.get('scan_path',{}).get('value'), d['e2e']['value'], d['n_gpus'], d['clocks'], d['config'].get('host_numa_node'))"
[ "${3:-extra}" = "noextra" ] || timeout 900 $TR bench_extra.py --only cfg3,cfg4 2>&1 | grep '^{' > gpurun_out/bench_extra_${N}gpu_$TAG.jsonl; cut -c1-260 gpurun_out/bench_extra_${N}gpu_$TAG.jsonl
