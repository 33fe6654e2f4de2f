"""CPU: the host-side logic of the Reasoner mirror (kolibrie_b200/engine.py: dictionary, rule compilation, store synchronisation,
incremental inference bookkeeping) with the oracle standing behind the Context method names (tests/oracle_ctx.py). The same flows run
against the device in tests/test_gpu_datalog.py and tests/test_gpu_sharded_fixpoint.py."""
import numpy as np
import pytest

from kolibrie_b200.engine import Constant, Reasoner, Rule, Variable
from tests import helpers as H
from tests.oracle_ctx import OracleCtx

FC = H.load("datalog_fc.json")["cases"]


@pytest.mark.parametrize("case", FC, ids=[x["name"] for x in FC])
def test_fc_through_the_reasoner_mirror(case):
    """datalog/tests/reasoning_tests.rs:28-404 as the reference writes them: add_abox_triple / add_rule / infer / query_abox"""
    d, facts, rules = H.build_fc_case(case)
    r = Reasoner(OracleCtx())
    for s, p, o in case["facts"]:
        r.add_abox_triple(s, p, o)
    for name in case["encode_after"]:
        r.dictionary.encode(name)
    for rule in rules:
        r.add_rule(rule)
    new_facts = r.infer_new_facts_semi_naive()
    for t in case["present"]:
        assert len(r.query_abox(*t)) > 0, f"{t} should be derivable"
    for t in case["absent"]:
        assert len(r.query_abox(*t)) == 0, f"{t} must not be derivable"
    if case.get("expect_empty"):
        assert new_facts == []
    if case.get("idempotent"):
        assert r.infer_new_facts_semi_naive() == []
        assert r.infer_new_facts_incremental() == []


def ancestors(triples):
    r = Reasoner(OracleCtx())
    for t in triples:
        r.add_abox_triple(*t)
    anc, par = r.dictionary.encode("ancestor"), r.dictionary.encode("parent")
    r.add_rule(Rule([(Variable("X"), Constant(par), Variable("Y"))], [(Variable("X"), Constant(anc), Variable("Y"))]))
    r.add_rule(Rule([(Variable("X"), Constant(anc), Variable("Y")), (Variable("Y"), Constant(anc), Variable("Z"))], [(Variable("X"), Constant(anc), Variable("Z"))]))
    return r


def decoded(r, ts):
    return {tuple(r.dictionary.id_to_string[x] for x in t) for t in ts}


FIRST = [(f"p{i}", "parent", f"p{i + 1}") for i in range(12)] + [("x", "likes", "y")]
LATER = [("p12", "parent", "p13"), ("q", "parent", "p0"), ("x", "likes", "z")]


def test_incremental_inference_equals_starting_over():
    r = ancestors(FIRST)
    one = r.infer_new_facts_semi_naive()
    for t in LATER:
        r.add_abox_triple(*t)
    two = r.infer_new_facts_incremental()
    assert len(two) == len(set(two)) and not (set(two) & set(one))
    fresh = ancestors(FIRST + LATER)
    assert decoded(r, one) | decoded(r, two) == decoded(fresh, fresh.infer_new_facts_semi_naive())
    assert len(r.query_abox("q", "ancestor", None)) == 14 and len(r.query_abox("x", "likes", None)) == 2
    assert r.infer_new_facts_incremental() == []
    # every stored triple once: base + inferred, nothing doubled by the seed path
    rows = r.ctx._rows()
    assert len(np.unique(rows, axis=0)) == len(rows) == len(r._facts)


def test_incremental_inference_after_the_store_was_reloaded_or_the_rules_changed():
    """a query between add_abox_triple and the inference reloads the device store WITH the added triples: there is no closed store plus
    a seed any more, the incremental call must start over (and still return exactly the new facts); so must a call after add_rule"""
    r = ancestors(FIRST)
    one = r.infer_new_facts_semi_naive()
    for t in LATER:
        r.add_abox_triple(*t)
    assert len(r.query_abox("x", "likes", None)) == 2  # reload
    two = r.infer_new_facts_incremental()
    fresh = ancestors(FIRST + LATER)
    assert decoded(r, one) | decoded(r, two) == decoded(fresh, fresh.infer_new_facts_semi_naive())
    assert not (set(two) & set(one))
    r.add_abox_triple("p13", "parent", "p14")
    desc = r.dictionary.encode("descendant")
    anc = r.dictionary.encode("ancestor")
    r.add_rule(Rule([(Variable("X"), Constant(anc), Variable("Y"))], [(Variable("Y"), Constant(desc), Variable("X"))]))
    three = r.infer_new_facts_incremental()
    n_anc = len(r.query_abox(None, "ancestor", None))
    assert len(r.query_abox(None, "descendant", None)) == n_anc and len([t for t in three if t[1] == desc]) == n_anc
