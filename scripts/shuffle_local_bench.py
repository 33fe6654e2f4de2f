"""One GPU, 8 'ranks' that are all local buffers: the SM-side cost of shuffle_scatter_kernel without NVLink (30 M rows x 2 columns,
as scripts/dist_shuffle_check.py times it across GPUs). With stores at HBM speed the floor is 0.48 GB / peak = 0.075 ms."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kolibrie_b200 import capi as c

n_parts = int(os.environ.get("PARTS", "8"))
nbig = 30_000_000
rng = np.random.default_rng(7)
ctx = c.Context(0)
X, Y = 0, 1
big = ctx.rel_from_host([X, Y], [rng.integers(0, 1 << 24, nbig).astype(np.uint32), rng.integers(0, 1 << 24, nbig).astype(np.uint32)])
cap = int(nbig / n_parts * 1.3)
bufs = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(2 * n_parts)]
cursors = torch.zeros(n_parts * 32, dtype=torch.int32, device="cuda")
peer_cols = [bufs[d * 2 + col].data_ptr() for d in range(n_parts) for col in range(2)]
peer_cur = [cursors.data_ptr() + 128 * d for d in range(n_parts)]
ctx.set_timing(True)
ks = []
for rep in range(6):
    cursors.zero_()
    torch.cuda.synchronize()
    ctx.get_stats(reset=True)
    ctx.shuffle_push(big, Y, n_parts, peer_cols, peer_cur, cap)
    ctx.synchronize()
    ks.append(ctx.get_stats(reset=True)["other_ms"])
got = cursors.view(n_parts, 32)[:, 0].cpu().numpy()
assert int(got.sum()) == nbig, got
print(f"shuffle_push, {n_parts} local destinations, {nbig} rows x 2 columns: kernel {min(ks[1:]):.3f} ms ({8 * nbig / (min(ks[1:]) * 1e-3) / 1e9:.0f} GB/s of rows)")
