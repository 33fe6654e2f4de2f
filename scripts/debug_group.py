import numpy as np, sys
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c, datagen
ctx = c.Context(0)
for n in (64, 20000):
    keys = (np.arange(n) % 3).astype(np.uint32) + 5
    vals = np.arange(n, dtype=np.uint32)
    ctx.dict_numeric_load(np.arange(n + 10, dtype=np.float64), np.ones(n + 10, dtype=np.uint8))
    rel = ctx.rel_from_host([1, 2], [keys, vals])
    g = ctx.group_aggregate(rel, [1], [(c.AGG_COUNT, 0), (c.AGG_SUM, 2)])
    print(n, g)
d = datagen.employee_dataset(20000)
ctx.store_load(d.s, d.p, d.o); ctx.dict_numeric_load(d.num_or0, d.is_num)
js, pats, _ = datagen.employee_queries(d)["cfg3"]
rel = ctx.star_join(js, pats)
print(rel.info())
for gs in ([1], [2], [1, 4]):
    g = ctx.group_aggregate(rel, gs, [(c.AGG_COUNT, 0), (c.AGG_SUM, 2)])
    print(gs, len(g['counts']), g['counts'][:5], g['keys'][0][:5])
