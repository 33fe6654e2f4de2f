"""torchrun --nproc-per-node N: the cross-rank GROUP BY of a prepared plan (kb_plan_attach_peers: partial tables in peer memory,
device-side barrier, one merge kernel reading the peers' tables over NVLink) against the oracle run on the unsharded store, for
every aggregate kind; and the NCCL all-gather + kb_groups_merge variant beside it."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, datagen, dist as kd
from tests import oracle_api as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
E = 200_000
full = datagen.employee_dataset(E)
d = datagen.employee_shard(E, rank, world, prefix=50_000)
db = O.Db(full.s, full.p, full.o, full.num_or0, full.is_num)
js, pats, _ = datagen.employee_queries(full)["cfg3"]
_, pats2, filt2 = datagen.employee_queries(full)["cfg2"]
ctx = c.Context(local)
ctx.set_sharding(rank, world)
ctx.dict_numeric_load(full.num_or0, full.is_num)
ctx.store_load(d.s, d.p, d.o)
ctx.build_index()


def table(x):
    keys = np.stack(x["keys"], axis=1)
    order = np.lexsort(tuple(keys[:, k] for k in range(keys.shape[1] - 1, -1, -1)))
    return keys[order], x["counts"][order], [v[order] for v in x["values"]]


def same(g, w, what):
    gk, gc, gv = table(g)
    wk, wc, wv = table(w)
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc), what
    for a, b in zip(gv, wv):
        assert np.allclose(a, b, rtol=1e-12, atol=0), what


n_checked = 0
for pp, ff, gslot in ((pats, None, 1), (pats2, filt2, 1), (pats, None, 4)):
    orel = db.bgp(pp, ff)
    for aggs in ([(c.AGG_COUNT, 0)], [(c.AGG_SUM, 2)], [(c.AGG_AVG, 2)], [(c.AGG_MIN, 2)], [(c.AGG_MAX, 2)], []):
        w = db.group(orel, [gslot], aggs)
        plan = kd.attach_group_plan(ctx.prepare_star_join(js, pp, ff, group_slots=[gslot], aggs=aggs, ring=3))
        tickets = [plan.submit() for _ in range(3)]
        for i in range(4):  # 7 queries through a ring of 3
            g, n_local = plan.collect_groups(tickets.pop(0))
            same(g, w, (gslot, aggs, "peer", i))
            tickets.append(plan.submit())
        rows = 0
        while tickets:
            g, n_local = plan.collect_groups(tickets.pop(0))
            same(g, w, (gslot, aggs, "peer tail"))
        assert kd.sum_over_ranks(n_local, dev) == orel.n_rows
        plan.free()
        packed, n_local = ctx.star_join_aggregate_packed(js, pp, ff, [gslot], aggs)
        same(ctx.groups_merge(kd.allgather_groups(packed, dev)), w, (gslot, aggs, "nccl"))
        n_checked += 1
dist.barrier(device_ids=[local])
print(f"rank {rank}: {n_checked} GROUP BY shapes x (peer-memory merge inside the plan, NCCL all-gather + kb_groups_merge) == oracle on the unsharded store", flush=True)
ctx.close()
dist.destroy_process_group()
