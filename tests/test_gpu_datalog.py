"""GPU (-m gpu): Datalog fixpoint on the device vs the reference's fc_* known answers and vs the oracle, through the
reference-shaped Reasoner mirror (kolibrie_b200/engine.py)."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import datagen
from kolibrie_b200.engine import Reasoner, compile_rule
from tests import helpers as H
from tests import oracle_api as O

pytestmark = pytest.mark.gpu

FC = H.load("datalog_fc.json")["cases"]


def inferred(r, s, p, o):
    return len(r.query_abox(s, p, o)) > 0


@pytest.mark.parametrize("case", FC, ids=[x["name"] for x in FC])
def test_fc_like_the_reference_test(ctx, case):
    """reads like datalog/tests/reasoning_tests.rs: add_abox_triple / add_rule / infer_new_facts_semi_naive / query_abox"""
    d, facts, rules = H.build_fc_case(case)
    r = Reasoner(ctx)
    for s, p, o in case["facts"]:
        r.add_abox_triple(s, p, o)
    for name in case["encode_after"]:
        r.dictionary.encode(name)
    assert r.dictionary.id_to_string == d.id_to_string
    for rule in rules:
        r.add_rule(rule)
    new_facts = r.infer_new_facts_semi_naive()
    for t in case["present"]:
        assert inferred(r, *t), f"{t} should be derivable"
    for t in case["absent"]:
        assert not inferred(r, *t), f"{t} must not be derivable"
    if case.get("expect_empty"):
        assert new_facts == []
    if case.get("idempotent"):
        assert r.infer_new_facts_semi_naive() == [], "second inference pass derives nothing new"
        for t in case["exactly_once"]:
            assert len(r.query_abox(*t)) == 1
    # and bit-exact against the oracle, including the per-round delta sizes
    db = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table())
    want = db.fixpoint([compile_rule(x) for x in rules], c.SEMI_NAIVE)
    H.assert_same_bag(np.array(new_facts, dtype=np.uint32).reshape(-1, 3), want["facts"], case["name"])


@pytest.mark.parametrize("strategy", [c.SEMI_NAIVE, c.NAIVE, c.SEMI_NAIVE_PARALLEL])
def test_taxonomy_closure_vs_oracle(ctx, strategy):
    """config 4 shape at 1/1000 scale: R1 transitive subClassOf, R2 type propagation; rounds and per-round counts must match"""
    t = datagen.taxonomy_dataset(fanout=4, depth=5, n_instances=20000)
    rules = datagen.taxonomy_rules(t)
    ctx.store_load(t.s, t.p, t.o)
    rel, st = ctx.datalog_fixpoint(rules, strategy)
    want = O.Db(t.s, t.p, t.o).fixpoint(rules, strategy)
    assert want["status"] == 0
    H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], "closure")
    assert st.rounds == len(want["round_new"])
    assert [st.round_new[i] for i in range(st.rounds)] == want["round_new"]
    assert st.derivations == want["derivations"]
    # closed form: every class at depth k has k proper ancestors
    n_sc = sum((4 ** k) * k for k in range(6))
    sc = t.ids["rdfs:subClassOf"]
    got_sc = int((rel.to_numpy([0, 1, 2])[:, 1] == sc).sum()) + int((t.p == sc).sum())
    assert got_sc == n_sc
    # the store now holds base + inferred; a second run derives nothing
    rel2, st2 = ctx.datalog_fixpoint(rules, strategy)
    assert rel2.n_rows == 0 and st2.rounds == 0


def test_rule_filters_and_constants_quirks(ctx):
    """numeric rule filter (rules.rs:148-160) and quirk Q6: constants in subject/object premise positions are NOT enforced"""
    from kolibrie_b200.engine import Constant, FilterCondition, Rule, Variable

    r = Reasoner(ctx)
    for s, p, o in [("a", "age", "30"), ("b", "age", "17"), ("c", "age", "x"), ("a", "type", "Person"), ("b", "type", "Robot")]:
        r.add_abox_triple(s, p, o)
    age, adult, typ, human = (r.dictionary.encode(x) for x in ("age", "adult", "type", "human"))
    yes, person = r.dictionary.encode("yes"), r.dictionary.encode("Person")
    r.add_rule(Rule([(Variable("X"), Constant(age), Variable("A"))], [(Variable("X"), Constant(adult), Constant(yes))], [FilterCondition("A", ">=", "18")]))
    r.add_rule(Rule([(Variable("X"), Constant(typ), Constant(person))], [(Variable("X"), Constant(human), Constant(yes))]))
    got = set(r.infer_new_facts_semi_naive())
    d = r.dictionary
    facts = np.array(r._facts[:5], dtype=np.uint32)
    want = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table()).fixpoint([compile_rule(x) for x in r.rules])
    assert got == {tuple(int(v) for v in row) for row in want["facts"]}
    a, b = d.lookup("a"), d.lookup("b")
    assert (a, adult, yes) in got and (b, adult, yes) not in got
    assert (b, human, yes) in got, "quirk Q6: (?x type Person) matches every type triple in the reference"
    # the parallel strategy (semi_naive_parallel.rs) enforces the constant and skips the filter — same answer as the oracle's
    r2 = Reasoner(ctx)
    for t in [("a", "age", "30"), ("b", "age", "17"), ("c", "age", "x"), ("a", "type", "Person"), ("b", "type", "Robot")]:
        r2.add_abox_triple(*t)
    for x in ("age", "adult", "type", "human", "yes", "Person"):
        r2.dictionary.encode(x)
    assert r2.dictionary.id_to_string == d.id_to_string[: len(r2.dictionary.id_to_string)]
    r2.rules = list(r.rules)
    got2 = set(r2.infer_new_facts_semi_naive_parallel())
    want2 = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table()).fixpoint([compile_rule(x) for x in r.rules], c.SEMI_NAIVE_PARALLEL)
    assert got2 == {tuple(int(v) for v in row) for row in want2["facts"]}
    assert (b, human, yes) not in got2 and (b, adult, yes) in got2


def test_known_fact_set_grows_mid_launch(ctx):
    """a head predicate that starts empty and receives 4 M new facts from ONE launch: the known-fact set is sized for an expected
    doubling (at least 2^20 facts), the derive launch stops at its budget, the set is rebuilt larger with the facts appended so far
    and the launch is repeated — every fact must still come out exactly once"""
    n = 4_000_000
    s = np.arange(10, 10 + n, dtype=np.uint32)
    o = (s * np.uint32(7) + np.uint32(3)) % np.uint32(1 << 22) + np.uint32(10)
    p = np.full(n, 5, np.uint32)
    ctx.store_load(s, p, o)
    rule = {"premise": [c.pattern(c.V(0), c.K(5), c.V(1))], "conclusion": [c.pattern(c.V(1), c.K(6), c.V(0))], "filters": []}
    rel, st = ctx.datalog_fixpoint([rule], c.SEMI_NAIVE)
    assert st.inferred == n and st.rounds == 1 and st.derivations == n
    got = rel.to_numpy([0, 1, 2])
    assert (got[:, 1] == 6).all()
    order = np.argsort(got[:, 2], kind="stable")
    assert np.array_equal(got[order, 2], s) and np.array_equal(got[order, 0], o)
    # a second run over the store (which now holds the inferred facts) derives nothing
    rel2, st2 = ctx.datalog_fixpoint([rule], c.SEMI_NAIVE)
    assert st2.inferred == 0


def test_known_fact_set_skewed_subject(ctx):
    """every derived fact of one rule has the SAME subject (constant in the head) and many candidates are duplicates of each other:
    heavy skew on one half of the 64-bit set key; every fact must still be derived exactly once"""
    rng = np.random.default_rng(9)
    n = 300_000
    s = rng.integers(100, 5000, n).astype(np.uint32)
    o = rng.integers(10_000, 10_000 + 120_000, n).astype(np.uint32)  # ~36 % of the objects occur more than once
    tr = np.unique(np.stack([s, np.full(n, 5, np.uint32), o], axis=1), axis=0)
    ctx.store_load(tr[:, 0], tr[:, 1], tr[:, 2])
    rules = [{"premise": [c.pattern(c.V(0), c.K(5), c.V(1))], "conclusion": [c.pattern(c.K(7), c.K(6), c.V(1))], "filters": []},
             {"premise": [c.pattern(c.V(0), c.K(5), c.V(1))], "conclusion": [c.pattern(c.V(0), c.K(8), c.K(7))], "filters": []}]
    rel, st = ctx.datalog_fixpoint(rules, c.SEMI_NAIVE)
    got = rel.to_numpy([0, 1, 2])
    objs = np.unique(tr[:, 2])
    subs = np.unique(tr[:, 0])
    want = np.concatenate([np.stack([np.full(len(objs), 7, np.uint32), np.full(len(objs), 6, np.uint32), objs], axis=1),
                           np.stack([subs, np.full(len(subs), 8, np.uint32), np.full(len(subs), 7, np.uint32)], axis=1)])
    H.assert_same_bag(got, want, "constant-subject heads")
    assert st.derivations == 2 * len(tr) and st.rounds == 1
    db = O.Db(tr[:, 0], tr[:, 1], tr[:, 2])
    w = db.fixpoint(rules, c.SEMI_NAIVE)
    H.assert_same_bag(got, w["facts"], "vs oracle")


@pytest.mark.parametrize("slack", ["8192", "0"])
@pytest.mark.parametrize("strategy", [c.SEMI_NAIVE, c.NAIVE])
def test_partitioned_candidate_dedup_vs_oracle(monkeypatch, strategy, slack):
    """the radix-partitioned dedup (candidates partitioned by the high bits of their home slot, probed slice by slice) forced onto
    small inputs: tiny slices -> hundreds of partitions; slack 0 -> buckets overflow on skewed keys and take the direct path. Facts,
    rounds, per-round counts and the derivation count must equal the oracle's, exactly as with the direct kernel."""
    monkeypatch.setenv("KOLIBRIE_DERIVE_PART", "1")
    monkeypatch.setenv("KOLIBRIE_DERIVE_SLICE", "2048")     # 256 slots per slice
    monkeypatch.setenv("KOLIBRIE_DERIVE_MIN_ROWS", "16")
    monkeypatch.setenv("KOLIBRIE_DERIVE_SLACK", slack)
    cx = c.Context(0)
    try:
        for fanout, depth, n_inst in ((4, 5, 20000), (3, 6, 5000), (10, 3, 60000)):
            t = datagen.taxonomy_dataset(fanout=fanout, depth=depth, n_instances=n_inst)
            rules = datagen.taxonomy_rules(t)
            cx.store_load(t.s, t.p, t.o)
            rel, st = cx.datalog_fixpoint(rules, strategy)
            want = O.Db(t.s, t.p, t.o).fixpoint(rules, strategy)
            H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], "closure")
            assert st.rounds == len(want["round_new"])
            assert [int(st.round_new[i]) for i in range(st.rounds)] == [int(x) for x in want["round_new"]]
            assert int(st.derivations) == int(want["derivations"])
            assert cx.get_stats()["rows_built"] > 0, "the partitioned path was not taken"
        # a skewed head: every candidate of a rule has the same subject (one bucket takes nearly everything)
        rng = np.random.default_rng(11)
        n = 30000
        tr = np.unique(np.stack([rng.integers(10, 60, n), np.full(n, 1), rng.integers(100, 4000, n)], axis=1).astype(np.uint32), axis=0)
        rule = {"premise": [c.pattern(c.V(0), c.K(1), c.V(1))], "conclusion": [c.pattern(c.K(7), c.K(2), c.V(1)), c.pattern(c.V(1), c.K(3), c.K(7))]}
        cx.store_load(tr[:, 0], tr[:, 1], tr[:, 2])
        rel, st = cx.datalog_fixpoint([rule], strategy)
        want = O.Db(tr[:, 0], tr[:, 1], tr[:, 2]).fixpoint([rule], strategy)
        H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], "skewed heads")
        assert int(st.derivations) == int(want["derivations"])
    finally:
        cx.close()


def test_old_delta_scheme_same_facts_fewer_candidates(ctx):
    """KB_SEMI_NAIVE_OLD_DELTA (premises before the delta premise read only OLD facts): the inferred facts, the rounds and the new facts
    per round are those of semi_naive.rs:17-85 (= the oracle's); only the number of candidate derivations is smaller"""
    t = datagen.taxonomy_dataset(fanout=3, depth=5, n_instances=20000)
    rules = datagen.taxonomy_rules(t)
    ctx.store_load(t.s, t.p, t.o)
    rel, st = ctx.datalog_fixpoint(rules, c.SEMI_NAIVE_OLD_DELTA)
    want = O.Db(t.s, t.p, t.o).fixpoint(rules, c.SEMI_NAIVE)
    H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], "old/delta closure")
    assert st.rounds == len(want["round_new"]) and st.inferred == len(want["facts"])
    assert [int(x) for x in st.round_new[:st.rounds]] == [int(x) for x in want["round_new"][:st.rounds]]
    assert 0 < st.derivations < want["derivations"]
    # the reference's fixtures with joins inside a rule: sibling / uncle shapes (reasoning_tests.rs:107-135, 362-404)
    rng = np.random.default_rng(5)
    n = 3000
    PARENT, SIB, UNCLE = 1, 2, 3
    kids = np.arange(10, 10 + n, dtype=np.uint32)
    par = (10 + n + rng.integers(0, n // 3, size=n)).astype(np.uint32)
    s = np.concatenate([kids, par[: n // 2]]); o = np.concatenate([par, (10 + 2 * n + rng.integers(0, 50, size=n // 2)).astype(np.uint32)])
    p = np.full(len(s), PARENT, dtype=np.uint32)
    sib = {"premise": [c.pattern(c.V(0), c.K(PARENT), c.V(2)), c.pattern(c.V(1), c.K(PARENT), c.V(2))], "conclusion": [c.pattern(c.V(0), c.K(SIB), c.V(1))],
           "filters": [c.KbRuleFilter(0, c.CMP_NE, 1, 1, 0.0)]}
    uncle = {"premise": [c.pattern(c.V(0), c.K(SIB), c.V(1)), c.pattern(c.V(2), c.K(PARENT), c.V(1))], "conclusion": [c.pattern(c.V(0), c.K(UNCLE), c.V(2))],
             "filters": []}
    ctx.store_load(s, p, o)
    rel, st = ctx.datalog_fixpoint([sib, uncle], c.SEMI_NAIVE_OLD_DELTA)
    want = O.Db(s, p, o).fixpoint([sib, uncle], c.SEMI_NAIVE)
    H.assert_same_bag(rel.to_numpy([0, 1, 2]), want["facts"], "sibling / uncle")
    assert st.rounds == len(want["round_new"]) and st.derivations <= want["derivations"]
