#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r1f}
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== bench full"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_full_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline']['families'].items()}, d['roofline']['device_ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 28 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
for K in scan_kernel build_direct probe_fast; do
  F=$(echo $K | cut -d_ -f1)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${F}_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
done
ls gpurun_out | head -30
