"""Where a window slide's time goes (BASELINE configs[4] shape, device-resident slides): evict / append (profile + split) / query,
each bracketed by a synchronize. Run plain for host-side times, under `ncu --metrics gpu__time_duration.sum` for the kernels'."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kolibrie_b200 import capi as c, datagen

per, n_slides, width = 1_000_002, 16, 10
d = datagen.employee_dataset(per * n_slides // 6)
ctx = c.Context(0)
ctx.dict_numeric_load(d.num_or0, d.is_num)
js, pats, filt = datagen.employee_queries(d)["cfg2"]
ds, dp, do = (torch.from_numpy(x).cuda() for x in (d.s, d.p, d.o))
live = []
acc = {"evict": [], "append": [], "query": [], "noop_sync": []}
for t in range(n_slides):
    lo = t * per
    ctx.synchronize()
    t0 = time.perf_counter()
    if len(live) == width:
        ctx.store_evict(live.pop(0))
    ctx.synchronize()
    t1 = time.perf_counter()
    ctx.store_append_device(ds.data_ptr() + 4 * lo, dp.data_ptr() + 4 * lo, do.data_ptr() + 4 * lo, per, 100 + t)
    ctx.synchronize()
    t2 = time.perf_counter()
    live.append(100 + t)
    if t == 0:
        ctx.build_index()
    r = ctx.star_join(js, pats, filt)
    rows = r.n_rows
    r.free()
    ctx.synchronize()
    t3 = time.perf_counter()
    ctx.synchronize()
    t4 = time.perf_counter()
    if t >= width:
        acc["evict"].append(t1 - t0); acc["append"].append(t2 - t1); acc["query"].append(t3 - t2); acc["noop_sync"].append(t4 - t3)
for k, v in acc.items():
    print(f"{k:10s} median {np.median(v) * 1e3:.3f} ms  min {min(v) * 1e3:.3f} ms")
print("launches", ctx.get_stats()["kernel_launches"])
