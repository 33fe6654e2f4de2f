"""torchrun --nproc-per-node N: a 2-hop PATH join (?x p1 ?y . ?y p2 ?z) over a store sharded by hash(subject): the first pattern's
rows must be re-sharded by ?y (a non-subject key) before the local join — kb_partition + NCCL all-to-all (kolibrie_b200.dist).
Every rank checks its slice against the oracle run on the full store."""
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
print(f"rank {rank}: local rows {len(got)} ok={ok} global {total} want {len(want)}", flush=True)
assert ok and total == len(want)
dist.destroy_process_group()
