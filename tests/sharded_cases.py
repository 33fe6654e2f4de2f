"""Rule sets + base facts for the sharded / incremental Datalog tests (CPU and GPU), and the oracle-backed engine that stands in for
the device in the CPU tests. Test infrastructure only."""
import numpy as np

from kolibrie_b200 import capi as c
from kolibrie_b200 import datagen
from tests import oracle_api as O

SPREAD = 97  # term ids are multiplied by this so that small test datasets spread over several 1024-id shard blocks


def spread(a):
    return (np.asarray(a, np.uint64) * SPREAD).astype(np.uint32)


def taxonomy_case(fanout=3, depth=4, n_instances=3000):
    """config-4 shape: R1 joins subClassOf with itself on a non-subject key, R2 joins rdf:type (object) with subClassOf (subject)"""
    t = datagen.taxonomy_dataset(fanout=fanout, depth=depth, n_instances=n_instances)
    rows = np.stack([spread(t.s), t.p, spread(t.o)], axis=1)
    return rows, datagen.taxonomy_rules(t), None, t.ids


def family_case(n=1500, seed=5):
    """the sibling / uncle shapes of reasoning_tests.rs:107-135, 362-404: joins on the OBJECT of both premises, a rule filter, a second
    rule over the first one's head"""
    rng = np.random.default_rng(seed)
    PARENT, SIB, UNCLE = 1, 2, 3
    kids = np.arange(10, 10 + n, dtype=np.uint32)
    par = (10 + n + rng.integers(0, n // 3, size=n)).astype(np.uint32)
    s = np.concatenate([kids, par[: n // 2]])
    o = np.concatenate([par, (10 + 2 * n + rng.integers(0, 50, size=n // 2)).astype(np.uint32)])
    p = np.full(len(s), PARENT, dtype=np.uint32)
    rows = np.unique(np.stack([spread(s), p, spread(o)], axis=1), axis=0)
    sib = {"premise": [c.pattern(c.V(0), c.K(PARENT), c.V(2)), c.pattern(c.V(1), c.K(PARENT), c.V(2))], "conclusion": [c.pattern(c.V(0), c.K(SIB), c.V(1))],
           "filters": [c.KbRuleFilter(0, c.CMP_NE, 1, 1, 0.0)]}
    uncle = {"premise": [c.pattern(c.V(0), c.K(SIB), c.V(1)), c.pattern(c.V(2), c.K(PARENT), c.V(1))], "conclusion": [c.pattern(c.V(0), c.K(UNCLE), c.V(2))],
             "filters": []}
    return rows, [sib, uncle], None, {"parent": PARENT, "sibling": SIB, "uncle": UNCLE}


def chain_case(n=400, seed=9):
    """transitive closure of a sparse random graph: one predicate, many rounds (paths cross shard borders again and again)"""
    rng = np.random.default_rng(seed)
    E = 7
    a = rng.integers(0, n, size=n + n // 4)
    b = np.minimum(a + rng.integers(1, 6, size=len(a)), n - 1)  # forward edges only: a DAG, the closure stays small
    keep = a != b
    rows = np.unique(np.stack([spread(a[keep] + 5), np.full(int(keep.sum()), E), spread(b[keep] + 5)], axis=1).astype(np.uint32), axis=0)
    tc = {"premise": [c.pattern(c.V(0), c.K(E), c.V(1)), c.pattern(c.V(1), c.K(E), c.V(2))], "conclusion": [c.pattern(c.V(0), c.K(E), c.V(2))], "filters": []}
    return rows, [tc], None, {"edge": E}


def fuzz_case(seed):
    """the random rule sets of tests/test_gpu_fuzz.py (<= 3 premises, constants, filters, several heads), ids spread over the shards"""
    from tests.test_gpu_fuzz import random_rules

    tr, num, isn, rules = random_rules(seed)
    rows = np.stack([spread(tr[:, 0]), tr[:, 1], spread(tr[:, 2])], axis=1) if len(tr) else tr.reshape(0, 3)
    num2 = np.zeros(len(num) * SPREAD)
    isn2 = np.zeros(len(isn) * SPREAD, np.uint8)
    num2[::SPREAD], isn2[::SPREAD] = num, isn

    def sp(t):
        return t if t.is_var else c.K(int(t.value) * SPREAD)

    out = []
    for r in rules:
        prem = [c.pattern(sp(x.s), x.p, sp(x.o)) for x in r["premise"]]
        conc = [c.pattern(sp(x.s), x.p, sp(x.o)) for x in r["conclusion"]]
        out.append({"premise": prem, "conclusion": conc, "filters": r["filters"]})
    return rows, out, (num2, isn2), {}


def closure_of(rows, rules, numeric=None):
    """the oracle's inferred facts for base facts `rows` (status, facts)"""
    db = O.Db(rows[:, 0], rows[:, 1], rows[:, 2], *(numeric or ()))
    w = db.fixpoint(rules, c.SEMI_NAIVE)
    return w["status"], w["facts"]


def rule_predicates(rules):
    ps = set()
    for r in rules:
        for x in list(r["premise"]) + list(r["conclusion"]):
            if not x.p.is_var:
                ps.add(int(x.p.value))
    return ps


class OracleEngine:
    """dist.ShardedFixpoint's engine contract on the CPU oracle: load / closure / closure_seed. The seed call is the contract of
    kb_datalog_fixpoint_seed restated with full closures: accepted = the seed facts (of rule predicates) the store does not hold,
    inferred = closure(store + accepted) minus what is already there."""

    def __init__(self, rules, numeric=None):
        self.rules, self.numeric = rules, numeric
        self.rows = np.empty((0, 3), np.uint32)
        self.preds = rule_predicates(rules)

    def load(self, rows):
        self.rows = np.unique(np.asarray(rows, np.uint32).reshape(-1, 3), axis=0)

    def closure(self):
        st, inf = closure_of(self.rows, self.rules, self.numeric)
        assert st == 0
        self.rows = np.concatenate([self.rows, inf], axis=0)
        return inf

    def closure_seed(self, rows):
        rows = np.asarray(rows, np.uint32).reshape(-1, 3)
        rows = rows[np.isin(rows[:, 1], np.fromiter(self.preds, np.uint32, len(self.preds)))] if len(rows) else rows
        have = set(map(tuple, self.rows.tolist()))
        acc = [t for t in dict.fromkeys(map(tuple, rows.tolist())) if t not in have]
        accepted = np.array(acc, np.uint32).reshape(-1, 3)
        self.rows = np.concatenate([self.rows, accepted], axis=0)
        if len(accepted) == 0:
            return accepted, np.empty((0, 3), np.uint32)
        st, inf = closure_of(self.rows, self.rules, self.numeric)
        assert st == 0
        self.rows = np.concatenate([self.rows, inf], axis=0)
        return accepted, inf
