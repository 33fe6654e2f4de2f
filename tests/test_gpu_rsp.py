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
