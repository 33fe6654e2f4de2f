#!/bin/bash
set -u
mkdir -p gpurun_out
for L in libkolibrie_b200.so libkb_t128.so libkb_t512.so; do
  echo "== $L"; KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/$L timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/b_var.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_var.json').read())
sp=d['scan_path']; print('scan',sp['value'],sp['ms_per_step'],{k:(round(v['ms'],4),round(v['frac'],3)) for k,v in sp['roofline']['families'].items()})
PY
done
KOLIBRIE_B200_LIB=$PWD/kolibrie_b200/libkb_t128.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rsp.py -m gpu -x -q 2>&1 | tail -2
