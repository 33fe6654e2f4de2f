#!/bin/bash
python -m pytest tests/test_gpu_datalog.py tests/test_gpu_rsp.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python bench.py --config cfg4 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', l['seconds_all'], 'old_delta', l['old_delta_scheme']['seconds_all'])"
KOLIBRIE_TRACE=1 python scripts/datalog_trace.py 2>&1 | awk '/==== run 2/{p=1} p' | grep "ensure_set\|initial\|wall\|append\|split" | awk '$(NF-2)+0>0.25 || /wall/' | head
