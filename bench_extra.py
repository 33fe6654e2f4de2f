#!/usr/bin/env python
"""bench_extra.py — the other BASELINE.json configs on the device path (bench.py is the contract line; these are companions).
One JSON line per workload. Run single-GPU, or under torchrun for the sharded variants (cfg3 / cfg4 are weak-scaled per GPU).

  cfg1  employee 10K, 2-pattern BGP                                     bindings/s
  cfg3  100M triples, 4-pattern star + GROUP BY ?t COUNT                bindings/s (join rows aggregated per second)
  cfg4  Datalog 2-rule closure over the 50M-triple taxonomy shape       inferred facts/s (subClassOf broadcast, rdf:type sharded)
  cfg5  RSP window: 1M-triple slides, 10-slide window, 3-pattern BGP per slide   slides/s, bindings/s (upload of each slide included)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg1,cfg3,cfg4,cfg5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the big configs (1.0 = BASELINE sizes)")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on a bounded sample")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from kolibrie_b200 import capi as c
    from kolibrie_b200 import datagen
    from kolibrie_b200 import dist as kd

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = c.Context(local)
    only = set(args.only.split(","))

    def sync_all():
        ctx.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])

    def reduce(rows, dt):
        if world > 1:
            return kd.sum_over_ranks(rows, dev), kd.max_over_ranks(dt, dev)
        return rows, dt

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    if "cfg1" in only:
        d = datagen.employee_dataset(10000)
        ctx.dict_numeric_load(d.num_or0, d.is_num)
        ctx.store_load(d.s, d.p, d.o)
        ctx.build_index()  # build_all_indexes, once, before the queries
        js, pats, filt = datagen.employee_queries(d)["cfg1"]
        for _ in range(3):
            r = ctx.star_join(js, pats, filt)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps * 10):
            r = ctx.star_join(js, pats, filt)
            rows = r.n_rows
            r.free()
        sync_all()
        dt = (time.perf_counter() - t0) / (args.steps * 10)
        emit({"workload": "cfg1: employee 10K (60 000 triples), 2-pattern BGP", "value": rows / dt, "unit": "bindings/s", "ms_per_query": dt * 1e3, "rows": rows,
              "note": "latency-bound: one query = one kernel launch + one stream synchronisation"})

    if "cfg3" in only:
        E = int(16_666_667 * args.scale)
        d = datagen.employee_shard(E * world, rank, world)
        ctx.set_sharding(rank, world)
        ctx.dict_numeric_load(d.num_or0, d.is_num)
        ctx.store_load(d.s, d.p, d.o)
        ctx.build_index()
        js, pats, _ = datagen.employee_queries(d)["cfg3"]

        def step():  # kb_star_join_aggregate: with the index the GROUP BY is folded into the probe kernel, no joined row is written
            g, n = ctx.star_join_aggregate(js, pats, None, [1], [(c.AGG_COUNT, 0)])
            return n, g

        for _ in range(3):
            rows, g = step()
        sync_all()
        ctx.set_timing(True)
        ctx.get_stats(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rows, g = step()
        sync_all()
        dt = (time.perf_counter() - t0) / args.steps
        st3 = ctx.get_stats(reset=True)
        ctx.set_timing(False)
        rows_all, dt = reduce(rows, dt)
        assert int(g["counts"].sum()) == rows and len(g["counts"]) == 3
        emit({"workload": f"cfg3: {6 * E} triples per GPU x {world} GPU, 4-pattern star + GROUP BY ?t COUNT", "value": rows_all / dt, "unit": "bindings/s",
              "ms_per_query": dt * 1e3, "groups": 3, "n_gpus": world, "scaling": "weak",
              "device_ms_per_query": {"join+group (one kernel)": st3["probe_ms"] / args.steps, "group": st3["group_ms"] / args.steps}})

    if "cfg4" in only:
        n_inst = int(48_888_890 * args.scale)
        t = datagen.taxonomy_dataset(10, 6, n_inst, seed=43, first_instance=rank * n_inst)  # every rank: the whole class tree + its slice of the instances
        rules = datagen.taxonomy_rules(t)
        kd.check_broadcast_plan(rules, [t.ids["rdfs:subClassOf"]]) if world > 1 else None
        times = []
        for rep in range(4):  # one warm-up closure (memory pool, code paths), then three timed ones, each on a freshly loaded store
            ctx.store_load(t.s, t.p, t.o)
            sync_all()
            t0 = time.perf_counter()
            rel, st = ctx.datalog_fixpoint(rules)
            sync_all()
            if rep:
                times.append(time.perf_counter() - t0)
            if rep < 3:
                rel.free()
        dt = sorted(times)[1]  # median: the stream-ordered pool occasionally has to map fresh memory (tens of ms per GB)
        inferred = int(st.inferred)
        sc_new = 5432100 if args.scale >= 1e-9 else 0
        if world > 1:  # subClassOf closure is derived identically on every rank: count it once
            inferred_all = kd.sum_over_ranks(inferred - (sc_new if rank else 0), dev)
            dt = kd.max_over_ranks(dt, dev)
        else:
            inferred_all = inferred
        emit({"workload": f"cfg4: Datalog R1+R2 over {len(t.s)} triples per GPU x {world} GPU (10-ary class tree depth 6 + rdf:type facts)", "value": inferred_all / dt,
              "unit": "inferred facts/s", "seconds": dt, "seconds_all": [round(x, 4) for x in times], "inferred": inferred_all, "rounds": int(st.rounds), "n_gpus": world, "scaling": "weak",
              "plan": "subClassOf broadcast (replicated), rdf:type sharded by subject: no shuffle"})
        rel.free()
        if args.cpu and rank == 0:
            from tests import oracle_api as O

            ts = datagen.taxonomy_dataset(10, 4, 200_000, seed=43)
            t1 = time.perf_counter()
            w = O.Db(ts.s, ts.p, ts.o).fixpoint(datagen.taxonomy_rules(ts))
            dtc = time.perf_counter() - t1
            emit({"workload": "cfg4 CPU oracle (reference's semi-naive algorithm restated), 10-ary tree depth 4 + 200 000 type facts", "value": len(w["facts"]) / dtc,
                  "unit": "inferred facts/s", "seconds": dtc, "inferred": int(len(w["facts"]))})

    if "cfg5" in only:
        per = int(1_000_002 * args.scale) // 6 * 6
        n_slides, width = 14, 10
        d = datagen.employee_dataset(per * n_slides // 6)
        ctx.dict_numeric_load(d.num_or0, d.is_num)
        js, pats, filt = datagen.employee_queries(d)["cfg2"]
        hs, hp, ho = (torch.from_numpy(x).pin_memory().numpy() for x in (d.s, d.p, d.o))
        ctx.store_clear()
        live, rows_tot, t_acc, timed = [], 0, 0.0, 0
        for t in range(n_slides):
            lo, hi = t * per, (t + 1) * per
            sync_all()
            t0 = time.perf_counter()
            if len(live) == width:
                ctx.store_evict(live.pop(0))
            ctx.store_append(hs[lo:hi], hp[lo:hi], ho[lo:hi], tag=100 + t)
            live.append(100 + t)
            r = ctx.star_join(js, pats, filt)
            rows = r.n_rows
            r.free()
            ctx.synchronize()
            if t >= width - 1:  # steady state: a full window
                t_acc += time.perf_counter() - t0
                rows_tot += rows
                timed += 1
        emit({"workload": f"cfg5: RSP window {width} slides x {per} triples, slide = evict + append (H2D) + 3-pattern BGP + FILTER", "value": rows_tot / t_acc,
              "unit": "bindings/s", "slides_per_s": timed / t_acc, "ms_per_slide": t_acc / timed * 1e3, "triples_per_s_sustained": per * timed / t_acc,
              "note": "BASELINE stream rate is 1M triples/s: one slide per second; a slide costs ms_per_slide"})
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
