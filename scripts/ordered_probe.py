import os, sys, time
os.environ["KOLIBRIE_ORDERED"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import numpy as np
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c
ctx = c.Context(0)
n = 3000
s = np.arange(50, 50 + n, dtype=np.uint32)
tr = np.concatenate([np.stack([s, np.full(n, p, np.uint32), s + np.uint32(10000 * (p - 99))], axis=1) for p in (100, 101, 102)]).astype(np.uint32)
def T(name, fn):
    ctx.synchronize(); t0 = time.perf_counter(); r = fn(); ctx.synchronize(); print(f"{name:28s} {1e3*(time.perf_counter()-t0):9.3f} ms", flush=True); return r
T("store_load", lambda: ctx.store_load(tr[:, 0], tr[:, 1], tr[:, 2]))
T("build_index", lambda: ctx.build_index())
pats = [c.pattern(c.V(0), c.K(100), c.V(1)), c.pattern(c.V(0), c.K(101), c.V(2))]
for i in range(4):
    T(f"star_join #{i}", lambda: ctx.star_join(0, pats))
T("star_join_aggregate", lambda: ctx.star_join_aggregate(0, pats, None, [1], [(c.AGG_COUNT, 0)]))
ctx.set_use_index(False)
for i in range(2):
    T(f"star_join scan path #{i}", lambda: ctx.star_join(0, pats))
