"""cfg5 slide broken down: evict / append (H2D) / query, synchronised after each part"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c, datagen
per, n_slides, width = 1_000_002, 16, 10
d = datagen.employee_dataset(per * n_slides // 6)
ctx = c.Context(0)
ctx.dict_numeric_load(d.num_or0, d.is_num)
js, pats, filt = datagen.employee_queries(d)["cfg2"]
hs, hp, ho = (torch.from_numpy(x).pin_memory().numpy() for x in (d.s, d.p, d.o))
ctx.store_clear()
live = []
acc = np.zeros(3); k = 0
for t in range(n_slides):
    lo, hi = t * per, (t + 1) * per
    ctx.synchronize(); t0 = time.perf_counter()
    if len(live) == width: ctx.store_evict(live.pop(0))
    ctx.synchronize(); t1 = time.perf_counter()
    ctx.store_append(hs[lo:hi], hp[lo:hi], ho[lo:hi], tag=100 + t); live.append(100 + t)
    ctx.synchronize(); t2 = time.perf_counter()
    r = ctx.star_join(js, pats, filt); rows = r.n_rows; r.free()
    ctx.synchronize(); t3 = time.perf_counter()
    if t >= width: acc += [t1 - t0, t2 - t1, t3 - t2]; k += 1
print("ms per slide: evict %.3f append %.3f query %.3f" % tuple(acc / k * 1e3), "rows", rows)
ctx.set_timing(True); ctx.get_stats(reset=True)
r = ctx.star_join(js, pats, filt); r.free()
print({k2: v for k2, v in ctx.get_stats().items() if k2.endswith("_ms") or k2.endswith("launches")})
