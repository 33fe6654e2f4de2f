#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_rsp.py -m gpu -x -q 2>&1 | tail -2
echo "== bench star"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_star1.json
python - <<PY
import json
d=json.loads(open('gpurun_out/b_star1.json').read())
sp=d['scan_path']; print('scan',sp['value'],sp['ms_per_step'],{k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_star -s 2 -c 1 -o gpurun_out/prof_scan_star_r2e -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
