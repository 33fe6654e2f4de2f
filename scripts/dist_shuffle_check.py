"""torchrun --nproc-per-node N: a 2-hop PATH join (?x p1 ?y . ?y p2 ?z) over a store sharded by hash(subject): the first pattern's
rows must be re-sharded by ?y (a non-subject key) before the local join. Both exchanges are checked: kb_partition + NCCL all-to-all
and the fused peer-memory kernel (kolibrie_b200.dist.PeerShuffle). Every rank checks its slice against the oracle run on the full
store; then both exchanges are timed on a large relation."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, datagen, dist as kd
from tests import oracle_api as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(7)
n = 400000
tr = np.unique(np.stack([rng.integers(0, 50000, n), rng.integers(100, 103, n), rng.integers(0, 50000, n)], axis=1).astype(np.uint32), axis=0)
s, p, o = kd.shard_triples(tr[:, 0], tr[:, 1], tr[:, 2], rank, world)
ctx = c.Context(local)
ctx.store_load(s, p, o)
X, Y, Z = 0, 1, 2
left, right = ctx.scan([c.pattern(c.V(X), c.K(100), c.V(Y)), c.pattern(c.V(Y), c.K(101), c.V(Z))])
# right is keyed by its SUBJECT ?y: already on the owner rank. left must move to the owner of ?y.
left_sh = kd.shuffle_relation(ctx, left, Y)
joined = ctx.hash_join(left_sh, right)
got = joined.to_numpy([X, Y, Z])
want = O.Db(tr[:, 0], tr[:, 1], tr[:, 2]).bgp([c.pattern(c.V(X), c.K(100), c.V(Y)), c.pattern(c.V(Y), c.K(101), c.V(Z))]).to_numpy([X, Y, Z])
mine = want[kd.shard_of(want[:, 1], world) == rank]
ok = np.array_equal(datagen.canonical_rows(got), datagen.canonical_rows(mine))
total = kd.sum_over_ranks(len(got), device=torch.device("cuda", local))
print(f"rank {rank}: NCCL exchange: local rows {len(got)} ok={ok} global {total} want {len(want)}", flush=True)
assert ok and total == len(want)

# ---- the same exchange fused with its transfer: one kernel writes into the peers' receive buffers over NVLink
import time
ps = kd.PeerShuffle(ctx, n_cols=2, capacity_rows=40_000_000)
H = datagen.canonical_rows
for name, fn in (("push (receiver-owned cursors)", lambda r: ps.shuffle(r, Y, copy=True)), ("planned (count matrix)", lambda r: ps.shuffle_planned(r, Y))):
    left_p2p = fn(left)
    ok2 = np.array_equal(H(left_p2p.to_numpy(sorted(left_p2p.slots))), H(left_sh.to_numpy(sorted(left_sh.slots))))
    got2 = ctx.hash_join(left_p2p, right).to_numpy([X, Y, Z])
    ok3 = np.array_equal(H(got2), H(mine))
    print(f"rank {rank}: peer-memory exchange [{name}]: same rows as NCCL {ok2}, join ok {ok3}", flush=True)
    assert ok2 and ok3

# ---- timing on a large 2-column relation (rows per rank)
nbig = 30_000_000
big = ctx.rel_from_host([X, Y], [rng.integers(0, 1 << 24, nbig).astype(np.uint32), rng.integers(0, 1 << 24, nbig).astype(np.uint32)])
dev = torch.device("cuda", local)
ctx.set_timing(True)
for name, fn in (("partition + NCCL all_to_all", lambda: kd.shuffle_relation(ctx, big, Y)), ("fused peer-memory kernel, planned ranges", lambda: ps.shuffle_planned(big, Y)),
                 ("fused peer-memory kernel, push", lambda: ps.shuffle(big, Y))):
    ts, ks = [], []
    for rep in range(5):
        dist.barrier(device_ids=[local]); torch.cuda.synchronize(dev)
        ctx.get_stats(reset=True)
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(dev); dist.barrier(device_ids=[local])
        ts.append(time.perf_counter() - t0)
        ks.append(ctx.get_stats(reset=True)["other_ms"])
        m = r.n_rows
        r.free()
    t = kd.max_over_ranks(min(ts[1:]), device=dev)
    k = kd.max_over_ranks(min(ks[1:]), device=dev)
    if rank == 0:
        sent = 8 * nbig * (world - 1) / world
        print(f"{name}: {nbig} rows x 2 columns per rank, {world} ranks: whole exchange {t * 1e3:.3f} ms ({sent / t / 1e9:.1f} GB/s sent per rank over NVLink); "
              f"kernels alone {k:.3f} ms ({sent / (k * 1e-3) / 1e9:.1f} GB/s)", flush=True)
dist.destroy_process_group()
