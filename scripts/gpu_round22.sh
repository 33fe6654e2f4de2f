#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== index tests"; timeout 600 python -m pytest tests -m gpu -x -q -k "index" 2>&1 | tail -3
echo "== ordered index tests"; KOLIBRIE_ORDERED=1 timeout 600 python -m pytest tests -m gpu -x -q -k "index" 2>&1 | tail -3
for v in kolibrie_b200 kb_blocked; do
echo "== variant $v"; KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/lib$v.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/var_$v.json; python -c "
import json; d=json.load(open('gpurun_out/var_$v.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_index -s 1 -c 1 -o gpurun_out/prof_probe_index_r1m -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/prof_probe_index_r1m.ncu-rep
