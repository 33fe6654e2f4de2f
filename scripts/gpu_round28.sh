#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12; lscpu | grep -i "numa\|socket" | head
for f in "" "--no-numa"; do
echo "== bench $f"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu $f 2>&1 | tail -1 > gpurun_out/numa.json; python -c "
import json; d=json.load(open('gpurun_out/numa.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('host_numa_node'))"
done
echo "== rsp"; timeout 300 python scripts/rsp_breakdown.py 2>&1 | tail -3
