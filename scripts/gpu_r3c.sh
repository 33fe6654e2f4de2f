#!/bin/bash
for t in 0 1; do
KOLIBRIE_SET_TIGHT=$t python bench.py --config cfg4 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tight $t cfg4', l['seconds_all'], 'old_delta', l['old_delta_scheme']['seconds_all'])"
done
