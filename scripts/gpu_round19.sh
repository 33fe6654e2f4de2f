#!/bin/bash
# 2-GPU validation of the index path under torchrun + the full GPU suite
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== bench N=2"; timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_2gpu_r1k.json; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_r1k.json')); print(d['value'], d['ms_per_step'], d.get('scan_path',{}).get('value'), d['e2e']['value'], d['n_gpus'])"
echo "== ref arm N=2"; timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-300
echo "== extra N=2"; timeout 900 $TR bench_extra.py --only cfg3,cfg4 2>&1 | grep '^{' > gpurun_out/bench_extra_2gpu_r1k.jsonl; cut -c1-200 gpurun_out/bench_extra_2gpu_r1k.jsonl
