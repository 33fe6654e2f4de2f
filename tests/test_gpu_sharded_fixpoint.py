"""GPU (-m gpu): incremental materialisation (kb_datalog_fixpoint_seed) against the oracle, and the sharded fixpoint of
kolibrie_b200/dist.py with DEVICE engines — `world` contexts on one GPU, the exchange done in-process (the torch.distributed loop
around the same nodes is covered on CPU by tests/test_sharded_fixpoint.py and on a multi-GPU box by scripts/dist_datalog_check.py)."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import dist as kd
from tests import helpers as H
from tests import sharded_cases as S

pytestmark = pytest.mark.gpu

CASES = {"taxonomy": S.taxonomy_case, "family": S.family_case, "chain": S.chain_case}


def split_base_and_seed(rows, seed):
    """80 % of the base facts are loaded, the rest arrive later as the seed — together with repeats of facts the store holds already
    and repeats inside the seed itself (the sharded fixpoint delivers both)"""
    rng = np.random.default_rng(seed)
    late = rng.random(len(rows)) < 0.2
    base, fresh = rows[~late], rows[late]
    again = base[rng.integers(0, len(base), size=min(50, len(base)))] if len(base) else base
    sd = np.concatenate([fresh, again, fresh[: len(fresh) // 3]], axis=0)
    return base, sd[rng.permutation(len(sd))]


def check_seed_call(cx, rows, rules, numeric, strategy, tag):
    base, sd = split_base_and_seed(rows, 11)
    eng = S.OracleEngine(rules, numeric)
    eng.load(base)
    want1 = eng.closure()
    if numeric is not None:
        cx.dict_numeric_load(*numeric)
    cx.store_load(base[:, 0], base[:, 1], base[:, 2])
    rel, st = cx.datalog_fixpoint(rules, strategy)
    H.assert_same_bag(rel.to_numpy([0, 1, 2]), want1, f"{tag}: closure of the loaded facts")
    want_acc, want_inf = eng.closure_seed(sd)
    seed_rel = cx.rel_from_host([0, 1, 2], [np.ascontiguousarray(sd[:, k]) for k in range(3)])
    out, n_new, st2 = cx.datalog_fixpoint_seed(rules, seed_rel, strategy)
    got = out.to_numpy([0, 1, 2])
    assert n_new == len(want_acc) and st2.inferred == len(want_inf) and len(got) == n_new + st2.inferred, (tag, n_new, len(want_acc), st2.inferred, len(want_inf))
    H.assert_same_bag(got[:n_new], want_acc, f"{tag}: seed facts that were new")
    H.assert_same_bag(got[n_new:], want_inf, f"{tag}: facts inferred from the seed")
    assert sum(int(st2.round_new[i]) for i in range(min(st2.rounds, 64))) == st2.inferred or st2.rounds > 64
    # the store holds loaded + first closure + accepted + inferred, every fact once
    s, p, o = cx.store_download()
    H.assert_same_bag(np.stack([s, p, o], axis=1), eng.rows, f"{tag}: store after the incremental closure")
    # the same seed again: nothing is new; a fresh fixpoint over the store finds it closed
    out2, n_new2, st3 = cx.datalog_fixpoint_seed(rules, seed_rel, strategy)
    assert n_new2 == 0 and out2.n_rows == 0 and st3.inferred == 0
    rel3, st4 = cx.datalog_fixpoint(rules, strategy)
    assert rel3.n_rows == 0 and st4.rounds == 0, f"{tag}: closed under the rules"
    # an empty seed is a no-op
    empty = cx.rel_from_host([0, 1, 2], [np.empty(0, np.uint32)] * 3)
    out4, n_new4, st5 = cx.datalog_fixpoint_seed(rules, empty, strategy)
    assert n_new4 == 0 and out4.n_rows == 0 and st5.rounds == 0


@pytest.mark.parametrize("strategy", [c.SEMI_NAIVE, c.SEMI_NAIVE_OLD_DELTA])
@pytest.mark.parametrize("name", sorted(CASES))
def test_seed_closure_equals_fresh_closure(name, strategy):
    rows, rules, numeric, _ = CASES[name]()
    cx = c.Context(0)
    try:
        check_seed_call(cx, rows, rules, numeric, strategy, f"{name}/{strategy}")
    finally:
        cx.close()


def test_seed_closure_fuzzed_rule_sets():
    cx = c.Context(0)
    ran = 0
    try:
        for seed in range(40):
            rows, rules, numeric, _ = S.fuzz_case(seed)
            st, _ = S.closure_of(rows, rules, numeric) if len(rows) else (1, None)
            if st != 0 or len(rows) < 10:
                continue
            check_seed_call(cx, rows, rules, numeric, c.SEMI_NAIVE, f"fuzz {seed}")
            ran += 1
    finally:
        cx.close()
    assert ran >= 10


def test_seed_call_validates_its_arguments():
    rows, rules, _, _ = S.chain_case()
    cx = c.Context(0)
    try:
        cx.store_load(rows[:, 0], rows[:, 1], rows[:, 2])
        two = cx.rel_from_host([0, 1], [rows[:, 0], rows[:, 2]])
        with pytest.raises(c.KolibrieError) as e:
            cx.datalog_fixpoint_seed(rules, two)
        assert e.value.status == c.KB_E_INVALID
        wrong = cx.rel_from_host([0, 2, 1], [rows[:, 0], rows[:, 1], rows[:, 2]])
        with pytest.raises(c.KolibrieError):
            cx.datalog_fixpoint_seed(rules, wrong)
        n, _ = cx.store_size()
        assert n == len(rows), "a refused call leaves the store as it was"
    finally:
        cx.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_sharded_fixpoint_with_device_engines(name):
    """three ranks' worth of contexts on one GPU: placement, super-steps and seeds as on a multi-GPU box, closures on the device"""
    rows, rules, numeric, _ = CASES[name]()
    st, want = S.closure_of(rows, rules, numeric)
    world = 3
    cxs = [c.Context(0) for _ in range(world)]
    try:
        nodes = [kd.ShardedFixpoint(kd.DeviceFixpointEngine(cxs[r], rules), r, world, rules) for r in range(world)]
        own = kd.shard_of(rows[:, 0], world)
        parts = kd.run_sharded_fixpoint_local(nodes, [(rows[own == r, 0], rows[own == r, 1], rows[own == r, 2]) for r in range(world)])
        H.assert_same_bag(np.concatenate(parts, axis=0), want, f"{name}: union over the ranks == global closure, every fact once")
        for r, part in enumerate(parts):
            assert (kd.shard_of(part[:, 0], world) == r).all()
        assert max(n.steps for n in nodes) >= 2 and sum(n.sent_rows for n in nodes) > 0
    finally:
        for cx in cxs:
            cx.close()


def test_reasoner_mirror_incremental_inference():
    """reads like reasoning_tests.rs (fc transitive shape): infer, add_abox_triple, infer again — the second inference through the
    incremental entry returns what a Reasoner that starts over with all the triples returns beyond the first closure"""
    from kolibrie_b200.engine import Constant, Reasoner, Rule, Variable

    def build(triples):
        r = Reasoner(c.Context(0))
        for t in triples:
            r.add_abox_triple(*t)
        anc = r.dictionary.encode("ancestor")
        par = r.dictionary.encode("parent")
        r.add_rule(Rule([(Variable("X"), Constant(par), Variable("Y"))], [(Variable("X"), Constant(anc), Variable("Y"))]))
        r.add_rule(Rule([(Variable("X"), Constant(anc), Variable("Y")), (Variable("Y"), Constant(anc), Variable("Z"))], [(Variable("X"), Constant(anc), Variable("Z"))]))
        return r

    first = [(f"p{i}", "parent", f"p{i + 1}") for i in range(30)] + [("x", "likes", "y")]
    later = [("p30", "parent", "p31"), ("q", "parent", "p0"), ("x", "likes", "z"), ("p31", "parent", "p32")]
    r = build(first)
    one = set(r.infer_new_facts_semi_naive())
    for t in later:
        r.add_abox_triple(*t)
    two = r.infer_new_facts_incremental()
    assert len(two) == len(set(two)) and not (set(two) & one)
    fresh = build(first + later)
    # (the two dictionaries number the later terms differently: compare decoded triples)
    everything = set(fresh.infer_new_facts_semi_naive())
    dec = lambda rs, ts: {tuple(rs.dictionary.id_to_string[x] for x in t) for t in ts}
    assert dec(r, one) | dec(r, two) == dec(fresh, everything)
    assert len(r.query_abox("q", "ancestor", None)) == 33 and len(r.query_abox("x", "likes", None)) == 2
    assert r.infer_new_facts_incremental() == [], "nothing added since: nothing inferred"
    r.ctx.close()
    fresh.ctx.close()


@pytest.mark.parametrize("state", ["1", "0"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_a_run_of_seeds_with_and_without_the_kept_state(name, state, monkeypatch):
    """six seeds in a row (a growing window): with KOLIBRIE_FIX_STATE=1 (default) every call after the first continues from the relations
    and known-fact sets its predecessor left in the context; with 0 every call splits the store again. Same answers, call by call."""
    monkeypatch.setenv("KOLIBRIE_FIX_STATE", state)
    rows, rules, numeric, _ = CASES[name]()
    rng = np.random.default_rng(23)
    order = rng.permutation(len(rows))
    chunks = np.array_split(order, 8)
    base = rows[np.concatenate(chunks[:2])]
    eng = S.OracleEngine(rules, numeric)
    eng.load(base)
    want0 = eng.closure()
    cx = c.Context(0)
    try:
        cx.store_load(base[:, 0], base[:, 1], base[:, 2])
        rel, st = cx.datalog_fixpoint(rules)
        H.assert_same_bag(rel.to_numpy([0, 1, 2]), want0, f"{name}: first closure")
        for k in range(2, 8):
            sd = rows[chunks[k]]
            if k == 5:
                sd = np.concatenate([sd, rows[chunks[3]][:7]], axis=0)  # some triples the store has seen already
            want_acc, want_inf = eng.closure_seed(sd)
            seed_rel = cx.rel_from_host([0, 1, 2], [np.ascontiguousarray(sd[:, j]) for j in range(3)])
            out, n_new, st2 = cx.datalog_fixpoint_seed(rules, seed_rel)
            got = out.to_numpy([0, 1, 2])
            assert n_new == len(want_acc), (name, state, k)
            H.assert_same_bag(got[:n_new], want_acc, f"{name} state {state} seed {k}: accepted")
            H.assert_same_bag(got[n_new:], want_inf, f"{name} state {state} seed {k}: inferred")
            if k == 4:  # a store mutation between two seeds: the kept state is stale and must not be used
                extra = np.array([[123_456, 999_999, 7]], np.uint32)
                cx.store_append(extra[:, 0], extra[:, 1], extra[:, 2], 77)
                eng.rows = np.concatenate([eng.rows, extra], axis=0)
        s, p, o = cx.store_download()
        H.assert_same_bag(np.stack([s, p, o], axis=1), eng.rows, f"{name} state {state}: store at the end")
        st_full, want_full = S.closure_of(rows, rules, numeric)
        n_inferred_total = len(eng.rows) - len(rows) - 1
        assert n_inferred_total == len(want_full), "the closure reached seed by seed == the closure of all the triples at once"
    finally:
        cx.close()


def test_seed_closure_through_the_partitioned_dedup(monkeypatch):
    """the same comparison with the radix-partitioned candidate dedup forced onto small inputs (the seed's facts and the heads derived
    from them both go through derive_partition_kernel / derive_probe_kernel)"""
    monkeypatch.setenv("KOLIBRIE_DERIVE_PART", "1")
    monkeypatch.setenv("KOLIBRIE_DERIVE_SLICE", "2048")
    monkeypatch.setenv("KOLIBRIE_DERIVE_MIN_ROWS", "16")
    cx = c.Context(0)
    try:
        for name in sorted(CASES):
            rows, rules, numeric, _ = CASES[name]()
            check_seed_call(cx, rows, rules, numeric, c.SEMI_NAIVE, f"{name}/partitioned")
    finally:
        cx.close()
