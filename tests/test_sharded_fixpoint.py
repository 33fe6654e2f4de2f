"""CPU: the sharded semi-naive fixpoint with an exchange per super-step (kolibrie_b200/dist.py: exchange_plan, ShardedFixpoint,
run_sharded_fixpoint) — placement rules, the super-step loop in one process with oracle engines for world = 1..4, and the same loop
over torch.distributed (gloo, world size 2). The device engine of the same scheme is checked in tests/test_gpu_sharded_fixpoint.py."""
import os
import socket

import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import dist as kd
from tests import helpers as H
from tests import sharded_cases as S


def shards_of(rows, world):
    own = kd.shard_of(rows[:, 0], world)
    return [(rows[own == r, 0], rows[own == r, 1], rows[own == r, 2]) for r in range(world)]


def run_local(rows, rules, numeric, world, replicated=()):
    nodes = [kd.ShardedFixpoint(S.OracleEngine(rules, numeric), r, world, rules, replicated) for r in range(world)]
    parts = kd.run_sharded_fixpoint_local(nodes, shards_of(rows, world))
    return nodes, parts


def test_exchange_plan_placement():
    rows, rules, _, ids = S.taxonomy_case()
    sc, ty = ids["rdfs:subClassOf"], ids["rdf:type"]
    homed, rep = kd.exchange_plan(rules, [])
    # R1 joins (?a sc ?b),(?b sc ?c) on ?b = object of the first premise; R2 joins (?x type ?a),(?a sc ?b) on ?a = object of type
    assert homed == {sc, ty} and rep == set()
    homed, rep = kd.exchange_plan(rules, [sc])
    assert homed == set() and rep == {sc}, "with the TBox replicated every rule has one sharded premise: no second homes"
    rows, rules, _, ids = S.family_case()
    homed, _ = kd.exchange_plan(rules, [])
    assert ids["parent"] in homed, "sibling joins two parent facts on their OBJECT"
    V, K, pat = c.V, c.K, c.pattern
    three = {"premise": [pat(V(0), K(1), V(1)), pat(V(1), K(2), V(2)), pat(V(2), K(3), V(3))], "conclusion": [pat(V(0), K(4), V(3))], "filters": []}
    with pytest.raises(ValueError):
        kd.exchange_plan([three], [])
    with pytest.raises(ValueError):
        kd.exchange_plan([three], [2])  # the middle premise replicated: the two sharded ones are left without a common variable
    assert kd.exchange_plan([three], [1, 2])[0] == set(), "one sharded premise: evaluated wherever its facts live"
    with pytest.raises(ValueError):  # a sharded premise must not feed a replicated predicate (its facts would have to be broadcast)
        kd.exchange_plan([{"premise": three["premise"][:2], "conclusion": [pat(V(0), K(2), V(2))], "filters": []}], [2])
    apart = {"premise": [pat(V(0), K(1), V(1)), pat(V(2), K(2), V(3))], "conclusion": [pat(V(0), K(4), V(3))], "filters": []}
    with pytest.raises(ValueError):
        kd.exchange_plan([apart], [])
    both_subjects = {"premise": [pat(V(0), K(1), V(1)), pat(V(0), K(2), V(2))], "conclusion": [pat(V(1), K(4), V(2))], "filters": []}
    assert kd.exchange_plan([both_subjects], [])[0] == set(), "a join on the subject of both premises is local under subject sharding"


CASES = {"taxonomy": S.taxonomy_case, "family": S.family_case, "chain": S.chain_case}


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_super_steps_reach_the_global_closure(name, world):
    rows, rules, numeric, ids = CASES[name]()
    st, want = S.closure_of(rows, rules, numeric)
    assert st == 0 and len(want) > 0
    nodes, parts = run_local(rows, rules, numeric, world)
    got = np.concatenate(parts, axis=0)
    H.assert_same_bag(got, want, f"{name} world {world}: every inferred fact reported exactly once")
    for r, part in enumerate(parts):
        assert (kd.shard_of(part[:, 0], world) == r).all(), "a rank reports the facts whose subject it owns"
    if world > 1:
        assert sum(n.sent_rows for n in nodes) > 0 and max(n.steps for n in nodes) >= 2


def test_replicated_tbox_needs_no_second_homes_and_rank0_reports_it():
    rows, rules, numeric, ids = S.taxonomy_case()
    sc = ids["rdfs:subClassOf"]
    st, want = S.closure_of(rows, rules, numeric)
    for world in (2, 3):
        nodes, parts = run_local(rows, rules, numeric, world, replicated=[sc])
        H.assert_same_bag(np.concatenate(parts, axis=0), want, "replicated TBox")
        assert all((part[:, 1] != sc).all() for part in parts[1:]) and (parts[0][:, 1] == sc).any()
        n_sc_base = int((rows[:, 1] == sc).sum())
        own = kd.shard_of(rows[rows[:, 1] == sc, 0], world)
        # only the base TBox travels (each fact from its subject home to the world - 1 others); derived facts stay where they are
        assert sum(n.sent_rows for n in nodes) == n_sc_base * (world - 1), (sum(n.sent_rows for n in nodes), n_sc_base, np.bincount(own))


def _eligible_fuzz_seeds(limit=60):
    """the seeds of tests/test_gpu_fuzz.py's generator whose rule sets the placement scheme serves and the oracle accepts"""
    out = []
    for seed in range(limit):
        rows, rules, numeric, _ = S.fuzz_case(seed)
        try:
            kd.exchange_plan(rules, [])
        except ValueError:
            continue  # three sharded premises / two without a common variable
        if len(rows) and S.closure_of(rows, rules, numeric)[0] == 0:
            out.append(seed)
    return out


FUZZ = _eligible_fuzz_seeds()


def test_fuzz_generator_yields_enough_servable_rule_sets():
    assert len(FUZZ) >= 15, FUZZ


@pytest.mark.parametrize("seed", FUZZ)
def test_fuzzed_rule_sets(seed):
    rows, rules, numeric, _ = S.fuzz_case(seed)
    st, want = S.closure_of(rows, rules, numeric)
    for world in (2, 3):
        nodes, parts = run_local(rows, rules, numeric, world)
        H.assert_same_bag(np.concatenate(parts, axis=0), want, f"fuzz seed {seed} world {world}")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name in sorted(CASES):
            rows, rules, numeric, _ = CASES[name]()
            s, p, o = kd.shard_triples(rows[:, 0], rows[:, 1], rows[:, 2], rank, world)
            node = kd.ShardedFixpoint(S.OracleEngine(rules, numeric), rank, world, rules)
            mine = kd.run_sharded_fixpoint(node, s, p, o)
            assert (kd.shard_of(mine[:, 0], world) == rank).all()
            allf = kd._allgather_rows(mine)
            st, want = S.closure_of(rows, rules, numeric)
            H.assert_same_bag(allf, want, f"{name} over gloo")
            assert kd.sum_over_ranks(node.sent_rows) > 0
        # the config-4 way: TBox replicated, only its base facts travel; rank 0 reports the replicated predicate
        rows, rules, numeric, ids = S.taxonomy_case()
        sc = ids["rdfs:subClassOf"]
        s, p, o = kd.shard_triples(rows[:, 0], rows[:, 1], rows[:, 2], rank, world)
        node = kd.ShardedFixpoint(S.OracleEngine(rules, numeric), rank, world, rules, [sc])
        mine = kd.run_sharded_fixpoint(node, s, p, o)
        assert rank == 0 or (mine[:, 1] != sc).all()
        H.assert_same_bag(kd._allgather_rows(mine), S.closure_of(rows, rules, numeric)[1], "replicated TBox over gloo")
        assert kd.sum_over_ranks(node.sent_rows) == int((rows[:, 1] == sc).sum()) * (world - 1)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_over_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
