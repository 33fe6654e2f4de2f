#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "shuffle or partition or old_delta" 2>&1 | tail -3
python scripts/shuffle_local_bench.py
KOLIBRIE_SHUFFLE_DIRECT=0 python scripts/shuffle_local_bench.py
KOLIBRIE_SHUFFLE_THREADS=256 python scripts/shuffle_local_bench.py
PARTS=2 python scripts/shuffle_local_bench.py
PARTS=16 python scripts/shuffle_local_bench.py
ncu --set full --import-source on --clock-control none -k regex:shuffle_scatter -s 2 -c 1 -o gpurun_out/prof_shufflelocal_r2p -f python scripts/shuffle_local_bench.py > /dev/null 2>&1
ls -la gpurun_out/prof_shufflelocal_r2p.ncu-rep
