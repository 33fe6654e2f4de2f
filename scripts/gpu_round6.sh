#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (default=unordered)"; timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -12
echo "== pytest gpu (ordered)"; KOLIBRIE_ORDERED=1 timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6
for ORD in 0 1; do
echo "== bench full ORDERED=$ORD"; KOLIBRIE_ORDERED=$ORD timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_full_r1e_ord$ORD.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline']['families'].items()}, d['roofline']['device_ms_per_step'], d['e2e']['ms_per_step'])"
done
echo "== ncu scan"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1e python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
echo "== ncu probe"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_fast -s 3 -c 1 -o gpurun_out/prof_probe_r1e python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
