"""CPU: pins the oracle against every known-answer fixture the reference's own tests hold for the hot path, and cross-checks
its two evaluation modes against each other and against a brute-force nested loop."""
import itertools

import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import datagen
from kolibrie_b200.engine import Dictionary, compile_rule, rust_parse_f64
from tests import helpers as H
from tests import oracle_api as O

FC = H.load("datalog_fc.json")["cases"]


@pytest.mark.parametrize("case", FC, ids=[x["name"] for x in FC])
@pytest.mark.parametrize("strategy", [c.SEMI_NAIVE, c.NAIVE])
def test_fc_fixture(case, strategy):
    """datalog/tests/reasoning_tests.rs:28-404 replayed on the oracle."""
    d, facts, rules = H.build_fc_case(case)
    db = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table())
    res = db.fixpoint([compile_rule(r) for r in rules], strategy)
    assert res["status"] == 0
    inferred = {tuple(int(x) for x in r) for r in res["facts"]}
    base = {tuple(int(x) for x in r) for r in facts}
    allf = inferred | base
    for t in case["present"]:
        assert H.triple_ids(d, t) in allf, f"{t} should be derivable"
    for t in case["absent"]:
        assert H.triple_ids(d, t) not in allf, f"{t} must not be derivable"
    if case.get("expect_empty"):
        assert len(inferred) == 0
    assert len(inferred) == len(res["facts"]), "no duplicate inferred facts"
    assert not (inferred & base), "inferred facts are new facts only (infer_generic.rs:52)"
    if case.get("idempotent"):
        f2 = np.concatenate([facts, res["facts"]])
        db2 = O.Db(f2[:, 0], f2[:, 1], f2[:, 2], *d.numeric_table())
        again = db2.fixpoint([compile_rule(r) for r in rules], strategy)
        assert len(again["facts"]) == 0, "second inference pass derives nothing new"


PAR_OK = [x for x in FC if all(not r["filters"] and len(r["premise"]) <= 2 for r in x["rules"])]


@pytest.mark.parametrize("case", PAR_OK, ids=[x["name"] for x in PAR_OK])
def test_fc_fixture_parallel_strategy(case):
    """infer_new_facts_semi_naive_parallel (semi_naive_parallel.rs:11-176) has no test in the reference; on rule sets without
    filters and with at most two premises it must agree with the semi-naive strategy the fixtures pin."""
    d, facts, rules = H.build_fc_case(case)
    db = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table())
    a = db.fixpoint([compile_rule(r) for r in rules], c.SEMI_NAIVE)
    b = db.fixpoint([compile_rule(r) for r in rules], c.SEMI_NAIVE_PARALLEL)
    H.assert_same_bag(a["facts"], b["facts"], case["name"])


def test_parallel_strategy_enforces_constants_and_ignores_filters():
    """matches_rule_pattern (rules.rs:9-72) checks constants in subject/object positions — the hash-join strategies do not
    (quirk Q6) — and the parallel variant never evaluates rule.filters."""
    from kolibrie_b200.engine import Constant, FilterCondition, Rule, Variable

    d = Dictionary()
    facts = np.array([[d.encode(s), d.encode(p), d.encode(o)] for s, p, o in
                      [("a", "type", "Person"), ("b", "type", "Robot"), ("a", "parent", "P"), ("b", "parent", "P")]], dtype=np.uint32)
    typ, human, yes, person, parent, sib = (d.encode(x) for x in ("type", "human", "yes", "Person", "parent", "sibling"))
    rules = [Rule([(Variable("X"), Constant(typ), Constant(person))], [(Variable("X"), Constant(human), Constant(yes))]),
             Rule([(Variable("X"), Constant(parent), Variable("Z")), (Variable("Y"), Constant(parent), Variable("Z"))],
                  [(Variable("X"), Constant(sib), Variable("Y"))], [FilterCondition("X", "!=", "Y")])]
    db = O.Db(facts[:, 0], facts[:, 1], facts[:, 2], *d.numeric_table())
    par = {tuple(int(v) for v in r) for r in db.fixpoint([compile_rule(r) for r in rules], c.SEMI_NAIVE_PARALLEL)["facts"]}
    sn = {tuple(int(v) for v in r) for r in db.fixpoint([compile_rule(r) for r in rules], c.SEMI_NAIVE)["facts"]}
    a, b = d.lookup("a"), d.lookup("b")
    assert (a, human, yes) in par and (b, human, yes) not in par, "parallel: constant object enforced"
    assert (b, human, yes) in sn, "hash-join strategies: quirk Q6"
    assert (a, sib, a) in par and (a, sib, a) not in sn, "parallel: filters are not evaluated"


def test_integration_fixture_scan_counts():
    """kolibrie/tests/integration_test.rs:131-299 — Exact-filter scans are id-equality scans."""
    fx = H.load("integration_fixture.json")
    tr = np.array(fx["triples"], dtype=np.uint32)
    d = Dictionary()
    for t in fx["terms"]:
        d.encode(t)
    num, isn = d.numeric_table()
    db = O.Db(tr[:, 0], tr[:, 1], tr[:, 2], num, isn)
    ex = fx["expect"]
    S, P, Ob = c.V(0), c.V(1), c.V(2)
    assert db.scan(c.pattern(c.K(0), P, Ob)).n_rows == ex["subject==person1"]
    assert db.scan(c.pattern(S, c.K(d.lookup("ex:name")), Ob)).n_rows == ex["predicate==ex:name"] == ex["count(ex:name)"]
    assert db.scan(c.pattern(S, P, c.K(d.lookup("Jane Doe")))).n_rows == ex["object==Jane Doe"]
    # "numeric objects (parse i32) -> 3" (:163-168): the three numeric literals 30, 25, 2000
    assert int(isn[tr[:, 2]].sum()) == ex["numeric_objects"]
    subs = db.scan(c.pattern(S, c.K(d.lookup("ex:worksFor")), c.K(2))).to_numpy()[:, 0]
    assert sorted(subs.tolist()) == ex["worksFor_company1_subjects"]
    # :277-284 add a second email for person1
    extra = d.encode("john.smith@example.com")
    tr2 = np.concatenate([tr, np.array([[0, 5, extra]], dtype=np.uint32)])
    db2 = O.Db(tr2[:, 0], tr2[:, 1], tr2[:, 2])
    assert db2.scan(c.pattern(c.K(0), c.K(5), Ob)).n_rows == ex["person1_emails_after_add"]
    # the two joins the reference's test asserts the answers of (:286-299, :302-342): they pin the oracle's multi-pattern join + FILTER
    C_, E_, A_ = 0, 1, 2
    tech = [c.pattern(c.V(C_), c.K(d.lookup("ex:industry")), c.K(d.lookup("Technology"))), c.pattern(c.V(E_), c.K(d.lookup("ex:worksFor")), c.V(C_))]
    young = [c.pattern(c.V(C_), c.K(d.lookup("ex:name")), c.K(d.lookup("ACME Corp"))), c.pattern(c.V(E_), c.K(d.lookup("ex:worksFor")), c.V(C_)),
             c.pattern(c.V(E_), c.K(d.lookup("ex:age")), c.V(A_))]
    lt30 = [c.fop(c.F_CMP_NUM, slot=A_, cmp=c.CMP_LT, value=30.0)]
    for mode in (0, 1):  # columnar and faithful oracle modes
        assert sorted(db.bgp(tech, mode=mode).to_numpy([E_])[:, 0].tolist()) == ex["tech_employees"]
        assert sorted(db.bgp(young, lt30, mode=mode).to_numpy([E_])[:, 0].tolist()) == ex["young_acme_employees"]


def employee4():
    fx = H.load("employee4.json")
    d = Dictionary()
    rows = []
    # document order of the embedded RDF/XML (simple_select_synth_data.rs:16-52): name, title, workplaceHomepage, f/p, s/h, salary
    for iri, title, sal in fx["employees"]:
        for p, o in (("foaf:name", iri), ("foaf:title", title), ("foaf:workplaceHomepage", fx["workplace"]), ("ds:full_or_part_time", "F"),
                     ("ds:salary_or_hourly", "SALARY"), ("ds:annual_salary", sal)):
            rows.append((d.encode(iri), d.encode(p), d.encode(o)))
    return fx, d, np.array(rows, dtype=np.uint32)


@pytest.mark.parametrize("mode", [0, 1])
def test_employee4_queries(mode):
    fx, d, tr = employee4()
    db = O.Db(tr[:, 0], tr[:, 1], tr[:, 2], *d.numeric_table())
    e, s, w = 0, 1, 2
    sal = db.bgp([c.pattern(c.V(e), c.K(d.lookup("ds:annual_salary")), c.V(s))], mode=mode).to_numpy([e, s])
    got = sorted((d.decode(int(a)), d.decode(int(b))) for a, b in sal)
    assert got == sorted((x[0], x[2]) for x in fx["employees"])
    # benches/my_benchmark.rs:29-41 / cuda_example.rs:33-41: ?employee foaf:workplaceHomepage ?w . ?employee ds:annual_salary ?salary
    j = db.bgp([c.pattern(c.V(e), c.K(d.lookup("foaf:workplaceHomepage")), c.V(w)), c.pattern(c.V(e), c.K(d.lookup("ds:annual_salary")), c.V(s))],
               mode=mode).to_numpy([e, w, s])
    assert len(j) == fx["expect_rows"]
    assert sorted((d.decode(int(a)), d.decode(int(b)), d.decode(int(cc))) for a, b, cc in j) == sorted((x[0], fx["workplace"], x[2]) for x in fx["employees"])


def test_rust_parse_f64_oracle_and_host():
    fx = H.load("rust_parse_f64.json")
    for s, want in fx["accept"].items():
        for fn in (O.rust_parse_f64, rust_parse_f64):
            got = fn(s)
            assert got is not None, (s, fn)
            if want == "nan":
                assert got != got
            elif isinstance(want, str):
                assert got == float(want)
            else:
                assert got == want
    for s in fx["reject"]:
        assert O.rust_parse_f64(s) is None, s
        assert rust_parse_f64(s) is None, s


def brute_force(tr, pats):
    """nested-loop evaluator: every assignment of triples to patterns whose variable bindings are consistent"""
    rows = []
    slots = []
    for p in pats:
        for t in (p.s, p.p, p.o):
            if t.is_var and t.value not in slots:
                slots.append(t.value)
    for combo in itertools.product(range(len(tr)), repeat=len(pats)):
        b = {}
        ok = True
        for p, i in zip(pats, combo):
            for t, v in zip((p.s, p.p, p.o), tr[i]):
                if t.is_var:
                    if b.setdefault(t.value, int(v)) != int(v):
                        ok = False
                elif t.value != int(v):
                    ok = False
            if not ok:
                break
        if ok:
            rows.append([b[s] for s in slots])
    return np.array(rows, dtype=np.uint32).reshape(-1, len(slots)), slots


@pytest.mark.parametrize("seed", range(6))
def test_columnar_vs_faithful_vs_bruteforce(seed):
    """multi-pattern BGP rows are pinned by no reference test: cross-check the oracle's modes and a brute-force evaluator."""
    rng = np.random.default_rng(seed)
    n = 60
    tr = np.unique(np.stack([rng.integers(0, 12, n), rng.integers(20, 24, n), rng.integers(0, 12, n)], axis=1).astype(np.uint32), axis=0)
    db = O.Db(tr[:, 0], tr[:, 1], tr[:, 2])
    x, y, z, w = 0, 1, 2, 3
    shapes = [
        [c.pattern(c.V(x), c.K(20), c.V(y)), c.pattern(c.V(x), c.K(21), c.V(z))],                                      # star on subject, 2 patterns
        [c.pattern(c.V(x), c.K(20), c.V(y)), c.pattern(c.V(x), c.K(21), c.V(z)), c.pattern(c.V(x), c.K(22), c.V(w))],  # 3-pattern star
        [c.pattern(c.V(x), c.K(20), c.V(y)), c.pattern(c.V(y), c.K(21), c.V(z))],                                      # path: object -> subject
        [c.pattern(c.V(x), c.V(y), c.K(3)), c.pattern(c.V(x), c.K(22), c.V(z))],                                       # variable predicate
    ]
    for pats in shapes:
        bf, slots = brute_force(tr, pats)
        col = db.bgp(pats, mode=0).to_numpy(slots)
        H.assert_same_bag(col, bf, "columnar vs brute force")
        if len(pats) >= 3 or len(pats) == 2:
            fa = db.bgp(pats, mode=1).to_numpy(slots)
            H.assert_same_bag(fa, bf, "faithful vs brute force")


def test_filter_semantics_and_group():
    d = datagen.employee_dataset(500)
    db = O.Db(d.s, d.p, d.o, d.num_or0, d.is_num)
    q = datagen.employee_queries(d)
    _, pats, filt = q["cfg2"]
    got = db.bgp(pats, filt).to_numpy([0, 1, 2, 3])
    want = int((d.salary_of_employee > 100000).sum())
    assert len(got) == want
    assert (d.num_or0[got[:, 2]] > 100000).all()
    assert np.array_equal(got[:, 0], got[:, 3]) or set(map(tuple, got[:, [0, 3]])) == {(a, a) for a in got[:, 0]}  # ?n has the subject's id
    # GROUP BY ?t COUNT (config 3 shape)
    _, pats3, _ = q["cfg3"]
    rel = db.bgp(pats3)
    g = db.group(rel, [1], [(c.AGG_COUNT, 0), (c.AGG_SUM, 2), (c.AGG_MIN, 2), (c.AGG_MAX, 2), (c.AGG_AVG, 2)])
    assert int(g["counts"].sum()) == d.n_employees
    for k, cnt, sm, mn, mx, av in zip(g["keys"][0], g["counts"], g["values"][1], g["values"][2], g["values"][3], g["values"][4]):
        t = [i for i, name in enumerate(datagen.POSITIONS) if d.ids.get(name) == int(k)][0]
        sal = d.salary_of_employee[d.title_of_employee == t]
        assert cnt == len(sal) and sm == sal.sum() and mn == sal.min() and mx == sal.max() and abs(av - sal.mean()) < 1e-9
