"""GPU (-m gpu): the RSP per-slide driver (config 5 shape) on the device store: each window firing evicts the oldest slide
(kb_store_evict), appends the new one (kb_store_append), re-materialises the rules and re-runs the plan — the sequence of
create_window_processor! (kolibrie/src/rsp_engine.rs:81-142) and SimpleR2R (kolibrie/src/rsp/simple_r2r.rs:95-142)."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import datagen
from tests import helpers as H
from tests import oracle_api as O

pytestmark = pytest.mark.gpu


def test_sliding_window_reevaluation(ctx):
    d = datagen.employee_dataset(12000)
    js, pats, filt = datagen.employee_queries(d)["cfg2"]
    ctx.dict_numeric_load(d.num_or0, d.is_num)
    ctx.store_clear()
    n_slides, width = 12, 4  # [RANGE 4 STEP 1] in units of one slide
    # the stream emits whole employees round-robin over the slides (timestamp = k // per_slide)
    per = d.n_triples // n_slides // 6 * 6
    live = []
    for t in range(n_slides):
        lo, hi = t * per, (t + 1) * per
        if len(live) == width:
            ctx.store_evict(live.pop(0))  # rsp_engine.rs:95-97 remove the previous window's triples
        ctx.store_append(d.s[lo:hi], d.p[lo:hi], d.o[lo:hi], tag=1000 + t)  # :101-104 add the current ones
        live.append(1000 + t)
        a, b = (t - len(live) + 1) * per, hi
        assert ctx.store_size() == (b - a, len(live))
        got = ctx.star_join(js, pats, filt)
        want = O.Db(d.s[a:b], d.p[a:b], d.o[a:b], d.num_or0, d.is_num).bgp(pats, filt)
        H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), f"slide {t}")


def test_window_with_reasoning(ctx):
    """reasoning-in-window (rsp_engine_test.rs:1203-1264 shape): materialise, query, drop the inferred facts, slide"""
    t = datagen.taxonomy_dataset(fanout=3, depth=3, n_instances=3000)
    rules = datagen.taxonomy_rules(t)
    sc = t.p == t.ids["rdfs:subClassOf"]
    tbox = (t.s[sc], t.p[sc], t.o[sc])
    abox = (t.s[~sc], t.p[~sc], t.o[~sc])
    ctx.store_clear()
    ctx.store_append(*tbox, tag=1)  # static background knowledge
    step = 1000
    for w in range(3):
        sl = slice(w * step, (w + 1) * step)
        ctx.store_append(abox[0][sl], abox[1][sl], abox[2][sl], tag=10 + w)
        if w > 0:
            ctx.store_evict(10 + w - 1)
        rel, st = ctx.datalog_fixpoint(rules)  # SimpleR2R::materialize = fresh inference over the live window
        s = np.concatenate([tbox[0], abox[0][sl]]); p = np.concatenate([tbox[1], abox[1][sl]]); o = np.concatenate([tbox[2], abox[2][sl]])
        want = O.Db(s, p, o).fixpoint(rules)
        H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], f"window {w}")
        # query over base + inferred: every instance of the window has its class chain
        q = ctx.scan([c.pattern(c.V(0), c.K(t.ids["rdf:type"]), c.V(1))])[0]
        assert q.n_rows == step + int((want["facts"][:, 1] == t.ids["rdf:type"]).sum())
        ctx.store_evict(c.KB_TAG_INFERRED)  # next firing re-materialises from scratch (simple_r2r.rs:103-128)


@pytest.mark.parametrize("q", ["cfg2", "cfg3", "star3"])
def test_window_slides_keep_the_index_path(ctx, q):
    """the store index is MAINTAINED across kb_store_append / kb_store_evict (one chunk per segment in every predicate slice, keys
    inserted into / cleared from the persistent tables in place): after every slide the star join still takes the one-kernel index
    path and still returns the oracle's rows for the live window; replays simple_r2r.rs:95-142 / rsp_engine.rs:94-104"""
    d = datagen.employee_dataset(14000)
    js, pats, filt = datagen.employee_queries(d)[q]
    ctx.dict_numeric_load(d.num_or0, d.is_num)
    ctx.store_clear()
    n_slides, width = 14, 5
    per = d.n_triples // n_slides // 6 * 6
    live = []
    # the window is built up from an indexed first slide; every later slide is an append (+ an eviction once the window is full)
    for t in range(n_slides):
        lo, hi = t * per, (t + 1) * per
        if len(live) == width:
            ctx.store_evict(live.pop(0))
        ctx.store_append(d.s[lo:hi], d.p[lo:hi], d.o[lo:hi], tag=500 + t)
        live.append(500 + t)
        if t == 0:
            ctx.build_index()  # SparqlDatabase::build_all_indexes once; never again
        a, b = (t - len(live) + 1) * per, hi
        n0 = ctx.get_stats()["index_joins"]
        got = ctx.star_join(js, pats, filt)
        assert ctx.get_stats()["index_joins"] == n0 + 1, f"slide {t}: the query left the index path"
        want = O.Db(d.s[a:b], d.p[a:b], d.o[a:b], d.num_or0, d.is_num).bgp(pats, filt)
        H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), f"slide {t}")
        # the fused GROUP BY and a prepared plan see the maintained index too
        if t in (3, 9):
            g, n_rows = ctx.star_join_aggregate(js, pats, filt, [1], [(c.AGG_COUNT, 0)])
            assert n_rows == want.n_rows and int(g["counts"].sum()) == want.n_rows
            plan = ctx.prepare_star_join(js, pats, filt, ring=2)
            assert plan.collect(plan.submit()) == want.n_rows
            plan.free()
    # a subject that re-appears in a later segment makes its predicates multi-valued: the tables go, the answers stay right
    ctx.store_append(d.s[a:a + 600], d.p[a:a + 600], d.o[a:a + 600], tag=999)
    got = ctx.star_join(js, pats, filt)
    s2 = np.concatenate([d.s[a:b], d.s[a:a + 600]]); p2 = np.concatenate([d.p[a:b], d.p[a:a + 600]]); o2 = np.concatenate([d.o[a:b], d.o[a:a + 600]])
    want = O.Db(s2, p2, o2, d.num_or0, d.is_num).bgp(pats, filt)
    H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), "duplicate subjects across segments")
    # evicting the duplicate segment leaves a correct (if table-less) index; a rebuild restores the fast path
    ctx.store_evict(999)
    got = ctx.star_join(js, pats, filt)
    want = O.Db(d.s[a:b], d.p[a:b], d.o[a:b], d.num_or0, d.is_num).bgp(pats, filt)
    H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), "after evicting the duplicates")
    ctx.build_index()
    n0 = ctx.get_stats()["index_joins"]
    got = ctx.star_join(js, pats, filt)
    assert ctx.get_stats()["index_joins"] == n0 + 1
    H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), "after the rebuild")


def test_slide_maintenance_corner_cases(ctx):
    """the one-kernel slide maintenance (segment_profile_kernel + segment_split_kernel) against the oracle where it has to leave its
    straight path: a slide that brings a NEW predicate, ids BELOW the persistent tables' first slot (rebuild), a subject repeated
    INSIDE a slide (the subject table goes), more than 8 predicates in one slide (batched scan path), a device-resident slide
    (kb_store_append_device), evictions in between — rsp_engine.rs:94-104 with arbitrary window contents"""
    import torch

    rng = np.random.default_rng(11)
    P = [50, 51, 52]
    js = 0
    pats = [c.pattern(c.V(0), c.K(50), c.V(1)), c.pattern(c.V(0), c.K(51), c.V(2)), c.pattern(c.V(0), c.K(52), c.V(3))]

    def slide(subj, preds=P, repeat=0):
        s = np.repeat(subj, len(preds)).astype(np.uint32)
        p = np.tile(np.array(preds, dtype=np.uint32), len(subj))
        o = (10_000 + rng.integers(0, 40, size=len(s))).astype(np.uint32)
        if repeat:
            s = np.concatenate([s, s[:repeat]]); p = np.concatenate([p, p[:repeat]]); o = np.concatenate([o, o[:repeat] + 1])
        return s, p, o

    ctx.store_clear()
    live = {}

    def check(label, expect_index=True):
        s = np.concatenate([v[0] for v in live.values()]); p = np.concatenate([v[1] for v in live.values()]); o = np.concatenate([v[2] for v in live.values()])
        n0 = ctx.get_stats()["index_joins"]
        got = ctx.star_join(js, pats, None)
        if expect_index:
            assert ctx.get_stats()["index_joins"] == n0 + 1, f"{label}: left the index path"
        want = O.Db(s, p, o).bgp(pats, None)
        H.assert_same_bag(got.to_numpy(sorted(got.slots)), want.to_numpy(sorted(want.slots)), label)
        for k, pt in enumerate(pats):  # single-pattern lookups through the maintained slices
            g1 = ctx.scan([pt])[0]
            w1 = O.Db(s, p, o).bgp([pt], None)
            H.assert_same_bag(g1.to_numpy(sorted(g1.slots)), w1.to_numpy(sorted(w1.slots)), f"{label} pattern {k}")

    def add(tag, sl, device=False):
        live[tag] = sl
        if device:
            d = [torch.from_numpy(x).cuda() for x in sl]
            ctx.store_append_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), len(sl[0]), tag)
        else:
            ctx.store_append(*sl, tag=tag)

    add(1, slide(np.arange(5000, 6000)))
    ctx.build_index()
    check("first slide")
    add(2, slide(np.arange(6000, 7000)))
    check("second slide")
    add(3, slide(np.arange(7000, 7500), preds=P + [53]))  # a predicate the index has not seen
    check("new predicate")
    g = ctx.scan([c.pattern(c.V(0), c.K(53), c.V(1))])[0]
    assert g.n_rows == 500
    add(4, slide(np.arange(100, 600)))  # ids below every table's first slot: the tables are rebuilt over all live chunks
    check("ids below the tables")
    ctx.store_evict(1); live.pop(1)
    check("after evicting the first slide")
    add(5, slide(np.arange(8000, 8400)), device=True)
    check("device-resident slide")
    add(6, slide(np.arange(9000, 9300), preds=list(range(50, 62))))  # 12 predicates: the batched scan path
    check("twelve predicates in a slide")
    ctx.store_evict(4); live.pop(4)
    ctx.store_evict(2); live.pop(2)
    check("after two more evictions")
    add(7, slide(np.arange(9500, 9800), repeat=30))  # subjects repeated inside the slide: predicate 50.. become multi-valued
    check("repeated subjects inside a slide", expect_index=False)
    ctx.store_evict(7); live.pop(7)
    check("after evicting the repeated subjects", expect_index=False)
    ctx.build_index()
    check("after the rebuild")
    # a window that slides far: the slices' id ranges follow the live chunks (the table walk does not cover the evicted past)
    tags = []
    for t in range(12):
        add(100 + t, slide(np.arange(20_000 + 1000 * t, 20_000 + 1000 * t + 700)))
        tags.append(100 + t)
        if len(tags) > 3:
            old = tags.pop(0)
            ctx.store_evict(old); live.pop(old)
    for tag in (3, 5, 6):
        ctx.store_evict(tag); live.pop(tag)
    check("long slide")
