#!/bin/bash
set -u
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null) nproc: $(nproc) affinity: $(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
python -c "
import sys; sys.path.insert(0,'.')
from tests import oracle_api as O
print('usable', O.usable_cpus(), 'oracle threads', O.num_threads())"
echo "== pytest gpu (durations)"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -10
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-420
