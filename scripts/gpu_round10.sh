#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest datalog"; timeout 600 python -m pytest tests/test_gpu_datalog.py tests/test_gpu_rsp.py tests/test_gpu_cpp_host.py -q -m gpu 2>&1 | tail -4
echo "== shuffle N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_shuffle_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
for N in 10000000 48888890; do echo "== datalog $N"; timeout 1200 python scripts/datalog_scale.py $N 2>&1 | tail -2 | tee -a gpurun_out/datalog_scale_b.jsonl | cut -c1-900; done
