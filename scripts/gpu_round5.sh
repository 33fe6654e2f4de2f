#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (ordered)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
for ORD in 1 0; do
echo "== bench full ORDERED=$ORD"; KOLIBRIE_ORDERED=$ORD timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_full_r1d_ord$ORD.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline']['families'].items()}, d['e2e']['ms_per_step'])"
done
echo "== ncu full scan ordered"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1d python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
echo "== ncu full scan unordered"; KOLIBRIE_ORDERED=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan_r1d_unord python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
