#!/bin/bash
for i in 1 2; do
python bench.py --no-cpu --no-e2e --no-adversarial 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=l['other_configs']
print('cache on: cfg4', o['cfg4']['seconds_all'], o['cfg4']['old_delta_scheme']['seconds_all'], 'cfg3', o['cfg3']['ms_per_step'], 'cfg5', o['cfg5']['ms_per_slide'], o['cfg5']['ms_per_slide_excl_h2d'], 'sync', l['sync_path']['ms_per_step'])"
done
KOLIBRIE_BUF_CACHE=0 python bench.py --no-cpu --no-e2e --no-adversarial 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=l['other_configs']
print('cache off: cfg4', o['cfg4']['seconds_all'], o['cfg4']['old_delta_scheme']['seconds_all'], 'cfg3', o['cfg3']['ms_per_step'], 'cfg5', o['cfg5']['ms_per_slide'], o['cfg5']['ms_per_slide_excl_h2d'], 'sync', l['sync_path']['ms_per_step'])"
KOLIBRIE_TRACE=1 python bench.py --config cfg4 --no-cpu 2>&1 >/dev/null | grep -i "kb trace" | awk '{ $1=""; print }' | sort | uniq -c | sort -rn | head -5
