"""Shared helpers for the test-suite: fixture loading and rule / pattern construction from string form."""
import json
import os

import numpy as np

from kolibrie_b200 import capi as c
from kolibrie_b200.engine import Dictionary, FilterCondition, Rule, Variable, Constant, compile_rule

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def build_fc_case(case):
    """Replays the setup of one reasoning_tests.rs fc_* test: add_abox_triple for each fact (encode s,p,o in order), then
    `enc` for the predicates the test encodes afterwards, then the rules. Returns (dictionary, facts[n,3], rules)."""
    d = Dictionary()
    facts = [(d.encode(s), d.encode(p), d.encode(o)) for s, p, o in case["facts"]]
    for name in case["encode_after"]:
        d.encode(name)

    def term(t):
        return Variable(t[1:]) if t.startswith("?") else Constant(d.encode(t))

    rules = []
    for r in case["rules"]:
        rules.append(Rule(premise=[tuple(term(t) for t in p) for p in r["premise"]], conclusion=[tuple(term(t) for t in p) for p in r["conclusion"]],
                          filters=[FilterCondition(*f) for f in r["filters"]]))
    return d, np.array(facts, dtype=np.uint32).reshape(-1, 3), rules


def triple_ids(d, t):
    return tuple(d.encode(x) for x in t)


def canon(a):
    from kolibrie_b200.datagen import canonical_rows

    return canonical_rows(a)


def assert_same_bag(a, b, msg=""):
    a, b = canon(a), canon(b)
    assert a.shape == b.shape, f"{msg} shape {a.shape} vs {b.shape}"
    assert np.array_equal(a, b), f"{msg} rows differ"
