#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== index tests with ring lib"; KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/libkb_ring.so timeout 600 python -m pytest tests -m gpu -x -q -k "index" 2>&1 | tail -2
KOLIBRIE_ORDERED=1 KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/libkb_ring.so timeout 600 python -m pytest tests -m gpu -x -q -k "index" 2>&1 | tail -2
for v in kolibrie_b200 kb_ring kolibrie_b200 kb_ring; do
echo "== variant $v"; KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/lib$v.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/var_$v.json; python -c "
import json; d=json.load(open('gpurun_out/var_$v.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
