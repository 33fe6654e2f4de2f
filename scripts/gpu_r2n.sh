#!/bin/bash
# multi-GPU validation + bench: N = $1
set -u
N=${1:-2}
TAG=${2:-r2n}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== group check N=$N"; timeout 600 $TR --master-port 29513 scripts/dist_group_check.py 2>&1 | grep "^rank\|Error\|error\|assert" | tail -10 | tee gpurun_out/dist_group_${N}gpu_${TAG}.txt
echo "== shuffle check N=$N"; timeout 900 $TR --master-port 29511 scripts/dist_shuffle_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -12 | tee gpurun_out/dist_shuffle_${N}gpu_${TAG}.txt
echo "== shuffle check N=$N, 2048-row tiles"; KOLIBRIE_SHUFFLE_THREADS=256 timeout 900 $TR --master-port 29514 scripts/dist_shuffle_check.py 2>&1 | grep "fused peer-memory kernel, push" | tee -a gpurun_out/dist_shuffle_${N}gpu_${TAG}.txt
echo "== bench reference arm N=$N"; timeout 600 $TR --master-port 29515 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu_${TAG}.json 2>/dev/null; tail -c 400 gpurun_out/bench_ref_${N}gpu_${TAG}.json
echo "== bench N=$N"; timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err; tail -c 1500 gpurun_out/bench_${N}gpu_${TAG}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${N}gpu_${TAG}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'ms',d['ms_per_step'],'host_us',d['details']['host_overhead_us_per_step'],'roof',d['roofline']['frac'])
    print('sync',d['sync_path']['value'],d['sync_path']['ms_per_step']); print('e2e',d['e2e']['value'])
    for k,v in d['multi_gpu'].items(): print(k,{a:b for a,b in v.items() if a not in('workload','exchange','parity','collective')})
except Exception as e: print('parse failed',e)
PY
