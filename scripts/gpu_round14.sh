#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "== bench_extra cfg3"; timeout 900 python bench_extra.py --only cfg3 2>&1 | grep "^{" | cut -c1-300
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
from kolibrie_b200 import capi as c, datagen
d = datagen.employee_dataset(16_666_667)
ctx = c.Context(0); ctx.store_load(d.s,d.p,d.o); ctx.dict_numeric_load(d.num_or0,d.is_num)
js,pats,_ = datagen.employee_queries(d)["cfg3"]
for i in range(3):
    r = ctx.star_join(js,pats); g = ctx.group_aggregate(r,[1],[(c.AGG_COUNT,0)])
ctx.get_stats(reset=True); ctx.set_timing(True)
for i in range(5):
    r = ctx.star_join(js,pats); g = ctx.group_aggregate(r,[1],[(c.AGG_COUNT,0),(c.AGG_SUM,2),(c.AGG_AVG,2)]); r.free()
st = ctx.get_stats()
print({k:(round(v/5,4) if isinstance(v,float) else v//5) for k,v in st.items()})
g2 = ctx.group_aggregate(ctx.star_join(js,pats),[2],[(c.AGG_COUNT,0)])
print("groups by salary", len(g2["counts"]))
PY
