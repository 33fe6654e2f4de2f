"""CPU, world_size 2, gloo: the host-side logic of the N>1 path — hash(subject) sharding and the join-key shuffle plumbing
(counts exchange + variable-size all-to-all). The device side of the same path (kb_partition) is checked in test_gpu_parity.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kolibrie_b200 import datagen
from kolibrie_b200 import dist as kd
from tests import oracle_api as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharding: this rank's shard of the global employee dataset, generated WITHOUT materialising the other shard
        E = 3000
        d = datagen.employee_shard(E, rank, world, prefix=700)
        full = datagen.employee_dataset(E)
        s, p, o = kd.shard_triples(full.s, full.p, full.o, rank, world)
        assert np.array_equal(d.s, s) and np.array_equal(d.p, p) and np.array_equal(d.o, o)
        assert (kd.shard_of(d.s, world) == rank).all()
        # 2. subject-star join needs no exchange: per-shard oracle results sum to the global result
        from kolibrie_b200 import capi as c  # struct constructors only
        js, pats, filt = datagen.employee_queries(full)["cfg2"]
        local = O.Db(d.s, d.p, d.o, full.num_or0, full.is_num).bgp(pats, filt).n_rows
        total = kd.sum_over_ranks(local)
        assert total == O.Db(full.s, full.p, full.o, full.num_or0, full.is_num).bgp(pats, filt).n_rows
        # 3. join-key shuffle: re-shard (subject, title) rows by the OBJECT (title) — a non-subject key
        rows = np.stack([d.s[1::6], d.o[1::6]], axis=1)  # ?e foaf:title ?t of this shard
        dest = kd.shard_of(rows[:, 1], world)
        order = np.argsort(dest, kind="stable")
        rows = rows[order]
        offs = [0] + list(np.cumsum(np.bincount(dest, minlength=world)))
        cols = [torch.from_numpy(rows[:, k].astype(np.int32)) for k in range(2)]
        recv = kd.all_to_all_relation(cols, [int(x) for x in offs])
        got = np.stack([t.numpy().astype(np.uint32) for t in recv], axis=1)
        assert (kd.shard_of(got[:, 1], world) == rank).all(), "every received row belongs to this rank"
        n_all = kd.sum_over_ranks(len(got))
        assert n_all == E, "the shuffle is a permutation of the global relation"
        mine = np.stack([full.s[1::6], full.o[1::6]], axis=1)
        mine = mine[kd.shard_of(mine[:, 1], world) == rank]
        assert np.array_equal(datagen.canonical_rows(got), datagen.canonical_rows(mine))
        assert kd.max_over_ranks(float(rank)) == world - 1
        # 4. Datalog (config 4 shape): broadcast subClassOf, keep rdf:type sharded by subject; union over ranks == global closure
        t = datagen.taxonomy_dataset(fanout=3, depth=4, n_instances=4000)
        rules = datagen.taxonomy_rules(t)
        ts, tp, to = kd.shard_triples(t.s, t.p, t.o, rank, world)
        local = lambda S, P, Ob: O.Db(S, P, Ob).fixpoint(rules)["facts"]
        mine = kd.datalog_fixpoint_sharded(local, ts, tp, to, rules, [t.ids["rdfs:subClassOf"]])
        allf = kd._allgather_rows(mine)
        want = O.Db(t.s, t.p, t.o).fixpoint(rules)["facts"]
        assert np.array_equal(datagen.canonical_rows(allf), datagen.canonical_rows(want)), "sharded closure == global closure"
        try:
            kd.check_broadcast_plan(rules, [])  # nothing replicated: R2 joins two sharded predicates on a non-subject key
            raise AssertionError("plan check should have refused")
        except ValueError:
            pass
        # 5. GROUP BY across ranks: per-shard partials in the kb_groups_pack layout, gathered in ONE collective (fixed slots with a
        #    length prefix; an oversized payload takes the second, sized collective), folded = the GROUP BY of the whole store
        gdb = O.Db(full.s, full.p, full.o, full.num_or0, full.is_num)
        js3, pats3, _ = datagen.employee_queries(full)["cfg3"]
        ldb = O.Db(d.s, d.p, d.o, full.num_or0, full.is_num)
        for gslot, slot_bytes in ((1, kd.GROUPS_SLOT_BYTES), (2, 4096)):  # by title: 3 groups; by salary: thousands (overflows a 4 KB slot)
            part = ldb.group(ldb.bgp(pats3), [gslot], [(c.AGG_AVG, 2)])
            sums = part["values"][0] * part["counts"]  # the device ships the SUM; the oracle hands back the average
            packed = kd.pack_groups_np(part["keys"], part["counts"], [sums], [c.AGG_AVG])
            parts = kd.allgather_bytes(packed, slot_bytes=slot_bytes)
            assert len(parts) == world and np.array_equal(parts[rank], packed)
            acc = {}
            for pb in parts:
                kk, cc, rr, kinds = kd.unpack_groups_np(pb)
                assert kinds == [c.AGG_AVG]
                for k, n_, v in zip(kk[:, 0], cc, rr[:, 0]):
                    a = acc.setdefault(int(k), [0, 0.0])
                    a[0] += int(n_)
                    a[1] += float(v)
            want = gdb.group(gdb.bgp(pats3), [gslot], [(c.AGG_AVG, 2)])
            assert sorted(acc) == sorted(int(k) for k in want["keys"][0])
            for k, n_, v in zip(want["keys"][0], want["counts"], want["values"][0]):
                assert acc[int(k)][0] == int(n_) and abs(acc[int(k)][1] / acc[int(k)][0] - v) <= 1e-9 * abs(v)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_shuffle_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_peer_shuffle_plan_ranges_are_disjoint_and_complete():
    """host side of the fused peer-memory shuffle: from the world x world matrix of row counts every sender derives where its rows
    start in every receiver's buffer; the ranges must tile each receive buffer exactly (source-rank order, no gaps, no overlap)"""
    from kolibrie_b200 import dist as kd

    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        m = rng.integers(0, 1000, (world, world))
        plans = [kd.shuffle_plan(m, r) for r in range(world)]
        for d in range(world):
            spans = sorted((plans[src][0][d], plans[src][0][d] + int(m[src, d])) for src in range(world))
            assert spans[0][0] == 0
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert spans[-1][1] == plans[d][1] == int(m[:, d].sum())
