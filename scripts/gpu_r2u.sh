#!/bin/bash
python scripts/dict_encode_bench.py 2>&1 | tail -3
KOLIBRIE_TRACE=1 python scripts/datalog_trace.py 2>&1 | awk '/==== run 2/{p=1} p' | grep -v "views 0.00\|ensure_set 0.00" | head -90
