"""torchrun --nproc-per-node N: the sharded Datalog fixpoint with an exchange per super-step (kolibrie_b200/dist.py: ShardedFixpoint,
DeviceFixpointEngine, run_sharded_fixpoint over NCCL) against the oracle's closure of the unsharded facts, for the rule shapes of
tests/sharded_cases.py, and a timing of the config-4 shape at 1/50 scale with NOTHING replicated (both rules join two sharded
predicates) beside the broadcast plan on the same data.
NOT RUN on a multi-GPU box in round 2 (the GPU budget was spent); the same nodes are exercised with device engines on one GPU by
tests/test_gpu_sharded_fixpoint.py and over gloo by tests/test_sharded_fixpoint.py."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, datagen, dist as kd
from tests import sharded_cases as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

for name, make in (("taxonomy", S.taxonomy_case), ("family", S.family_case), ("chain", S.chain_case)):
    rows, rules, numeric, _ = make()
    s, p, o = kd.shard_triples(rows[:, 0], rows[:, 1], rows[:, 2], rank, world)
    ctx = c.Context(local)
    node = kd.ShardedFixpoint(kd.DeviceFixpointEngine(ctx, rules), rank, world, rules)
    mine = kd.run_sharded_fixpoint(node, s, p, o, dev)
    allf = kd._allgather_rows(mine, dev)
    st, want = S.closure_of(rows, rules, numeric)
    ok = np.array_equal(datagen.canonical_rows(allf), datagen.canonical_rows(want))
    print(f"rank {rank}: {name}: {len(mine)} facts reported here, {len(allf)} in all, == oracle closure: {ok}; super-steps {node.steps}, rows sent {node.sent_rows}", flush=True)
    assert ok
    ctx.close()

# config-4 shape, 1 M triples: exchange scheme with nothing replicated vs broadcast plan with the TBox replicated
t = datagen.taxonomy_dataset(fanout=10, depth=4, n_instances=1_000_000)
rules = datagen.taxonomy_rules(t)
s, p, o = kd.shard_triples(t.s, t.p, t.o, rank, world)
for label, rep in (("exchange per super-step, nothing replicated", []), ("exchange scheme, TBox replicated", [t.ids["rdfs:subClassOf"]])):
    ctx = c.Context(local)
    node = kd.ShardedFixpoint(kd.DeviceFixpointEngine(ctx, rules), rank, world, rules, rep)
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    mine = kd.run_sharded_fixpoint(node, s, p, o, dev)
    torch.cuda.synchronize(); dt = kd.max_over_ranks(time.perf_counter() - t0, dev)
    total = kd.sum_over_ranks(len(mine), dev)
    if rank == 0:
        print(f"{label}: {total} inferred facts over {world} ranks in {dt * 1e3:.1f} ms wall (host-plumbed exchange included), "
              f"device time of the closures on rank 0 {node.engine.device_ms:.1f} ms, super-steps {node.steps}", flush=True)
    ctx.close()
dist.destroy_process_group()
