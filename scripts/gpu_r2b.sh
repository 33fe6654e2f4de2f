#!/bin/bash
# multi-GPU pass: N = $1
set -u
N=${1:-2}
mkdir -p gpurun_out
[ -n "${SKIP_SHUF:-}" ] || timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_shuffle_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -12 | tee gpurun_out/dist_shuffle_${N}gpu_r2b.txt
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_r2b.json 2> gpurun_out/bench_${N}gpu_r2b.err; tail -c 1500 gpurun_out/bench_${N}gpu_r2b.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${N}gpu_r2b.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'ms',d['ms_per_step'],'host_us',d['details']['host_overhead_us_per_step'],'roof',d['roofline']['frac'])
    print('sync',d['sync_path']['value'],d['sync_path']['ms_per_step']); print('e2e',d['e2e']['value'])
    for k,v in d['multi_gpu'].items(): print(k,{a:b for a,b in v.items() if a not in('workload','exchange','parity','collective')})
except Exception as e: print('parse failed',e)
PY
