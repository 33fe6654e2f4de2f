#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== ordered parity"; KOLIBRIE_ORDERED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rsp.py -m gpu -x -q 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/b40.json; python -c "
import json; d=json.load(open('gpurun_out/b40.json')); sp=d['scan_path']; print(d['value'], d['roofline']['ms_per_launch'], sp['value'], {k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})"
