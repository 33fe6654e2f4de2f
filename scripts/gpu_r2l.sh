#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_datalog.py -m gpu -x -q 2>&1 | tail -3
python scripts/shuffle_local_bench.py
PARTS=2 python scripts/shuffle_local_bench.py
ncu --set full --import-source on --clock-control none -k regex:shuffle_scatter -s 2 -c 1 -o gpurun_out/shuffle_local -f python scripts/shuffle_local_bench.py > /dev/null 2>&1
ls -la gpurun_out/shuffle_local.ncu-rep
timeout 300 python bench.py --config cfg4 2>/dev/null | tail -c 1500
