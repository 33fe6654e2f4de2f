import os, sys, time
os.environ["KOLIBRIE_ORDERED"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import numpy as np
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c
import tests.test_gpu_parity as tp
ctx = c.Context(0)
orig = ctx.star_join
def timed(*a, **k):
    ctx.synchronize(); t0 = time.perf_counter(); r = orig(*a, **k); ctx.synchronize()
    dt = time.perf_counter() - t0
    if dt > 0.01: print(f"  star_join took {dt*1e3:.1f} ms; pats={len(a[1])} filt={'y' if len(a)>2 and a[2] else 'n'}", flush=True)
    return r
ctx.star_join = timed
ob = ctx.build_index
def tb():
    t0 = time.perf_counter(); r = ob(); print(f"  build_index {1e3*(time.perf_counter()-t0):.1f} ms", flush=True); return r
ctx.build_index = tb
from tests import oracle_api as O
odb = O.Db.bgp
def tbgp(self, *a, **k):
    t0 = time.perf_counter(); r = odb(self, *a, **k); dt = time.perf_counter() - t0
    if dt > 0.05: print(f"  oracle bgp took {dt*1e3:.1f} ms", flush=True)
    return r
O.Db.bgp = tbgp
t0 = time.perf_counter()
tp.test_index_kernel_shapes(ctx, 1023)
print("test total", round(time.perf_counter() - t0, 2), "s")
