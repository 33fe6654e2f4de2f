"""where does a window slide's upload time go: torch pinned H2D vs kb_store_append for several sizes"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c
ctx = c.Context(0)
for n in (250_000, 1_000_002, 4_000_000, 16_000_000):
    a = [torch.randint(0, 1 << 20, (n,), dtype=torch.int32).pin_memory() for _ in range(3)]
    d = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for x, y in zip(a, d): y.copy_(x, non_blocking=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    npv = [x.numpy().view(np.uint32) for x in a]
    ts = []
    for rep in range(4):
        ctx.synchronize(); t2 = time.perf_counter()
        ctx.store_append(npv[0], npv[1], npv[2], tag=rep)
        ctx.synchronize(); t3 = time.perf_counter()
        ts.append(t3 - t2)
        ctx.store_evict(rep)
    print(f"n={n}: torch 3 pinned copies {1e3*(t1-t0):.3f} ms ({12*n/(t1-t0)/1e9:.1f} GB/s); kb_store_append {1e3*min(ts):.3f} ms ({12*n/min(ts)/1e9:.1f} GB/s)", flush=True)
