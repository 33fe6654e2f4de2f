"""The LEGACY executor's BGP + FILTER stage: the faithful Python restatement (tests/legacy_oracle.py) is
pinned against the reference's known answer for it (the 4-employee dataset of simple_select_synth_data.rs, which runs `execute_query`),
and the device path (kolibrie_b200.engine.LegacyExecutor -> kb_bgp_execute with KB_F_CMP_LEGACY) must return the same bags, quirks
included."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import engine as E
from tests import helpers as H
from tests import legacy_oracle as L

FOAF, DS = "http://xmlns.com/foaf/0.1/", "https://data.cityofchicago.org/resource/xzkq-xp2w/"


def employee4_triples():
    fx = H.load("employee4.json")
    tr = []
    for iri, title, sal in fx["employees"]:
        tr += [(iri, FOAF + "name", iri), (iri, FOAF + "title", title), (iri, FOAF + "workplaceHomepage", fx["workplace"]),
               (iri, DS + "full_or_part_time", "F"), (iri, DS + "salary_or_hourly", "SALARY"), (iri, DS + "annual_salary", sal)]
    return fx, tr


def bag(rows, keys):
    return sorted(tuple(r.get(k) for k in keys) for r in rows)


def test_rust_parse_i32():
    for s, v in (("0", 0), ("-5", -5), ("+7", 7), ("007", 7), ("2147483647", 2147483647), ("-2147483648", -2147483648)):
        assert E.rust_parse_i32(s) == v
    for s in ("", "+", "-", " 5", "5 ", "5.0", "1e3", "2147483648", "-2147483649", "0x10", "٣", "5\n", "--5"):
        assert E.rust_parse_i32(s) is None, s


def test_restatement_against_the_reference_known_answer():
    """simple_select_synth_data.rs:16-52 through execute_query: ?employee workplaceHomepage ?w . ?employee annual_salary ?salary -> 4 rows"""
    fx, tr = employee4_triples()
    rows = L.execute_bgp(tr, [("?employee", FOAF + "workplaceHomepage", "?workplaceHomepage"), ("?employee", DS + "annual_salary", "?salary")])
    assert len(rows) == fx["expect_rows"] == 4
    assert bag(rows, ["?employee", "?workplaceHomepage", "?salary"]) == sorted((x[0], fx["workplace"], x[2]) for x in fx["employees"])
    # FILTER(?salary > 80000): i32 comparison (apply_filters_simd)
    rows = L.execute_bgp(tr, [("?e", DS + "annual_salary", "?s")], [E.Comparison("?s", ">", "80000")])
    assert bag(rows, ["?e"]) == sorted((x[0],) for x in fx["employees"] if int(x[2]) > 80000)


QUERIES = [
    ("two patterns", [("?e", FOAF + "workplaceHomepage", "?w"), ("?e", DS + "annual_salary", "?s")], []),
    ("constant object is enforced", [("?e", FOAF + "title", "Developer"), ("?e", DS + "annual_salary", "?s")], []),
    ("constant subject is NOT enforced (a binding name)", [("http://example.org/employee1", FOAF + "title", "?t")], []),
    ("the same constant subject twice joins", [("http://example.org/employee1", FOAF + "title", "?t"), ("http://example.org/employee1", DS + "annual_salary", "?s")], []),
    ("variable predicate matches nothing", [("?e", "?p", "?o")], []),
    ("i32 filter", [("?e", DS + "annual_salary", "?s")], [E.Comparison("?s", ">=", "83504")]),
    ("i32 vs non-integer constant: string comparison, > never holds", [("?e", DS + "annual_salary", "?s")], [E.Comparison("?s", ">", "80000.0")]),
    ("string equality", [("?e", FOAF + "title", "?t")], [E.Comparison("?t", "=", "Manager")]),
    ("string inequality", [("?e", FOAF + "title", "?t")], [E.Comparison("?t", "!=", "Manager")]),
    ("ordering on strings never holds", [("?e", FOAF + "title", "?t")], [E.Comparison("?t", "<", "Manager")]),
    ("a constant the dictionary does not hold", [("?e", FOAF + "title", "?t")], [E.Comparison("?t", "!=", "Astronaut")]),
    ("nested comparisons parse f64", [("?e", DS + "annual_salary", "?s"), ("?e", FOAF + "title", "?t")],
     [E.And(E.Comparison("?s", ">", "70000.5"), E.Not(E.Comparison("?t", "=", "Manager")))]),
    ("OR of a numeric and a string test", [("?e", DS + "annual_salary", "?s"), ("?e", FOAF + "title", "?t")],
     [E.Or(E.Comparison("?s", "<", "70000"), E.Comparison("?t", "=", "Developer"))]),
    ("two top-level filters are a conjunction", [("?e", DS + "annual_salary", "?s")], [E.Comparison("?s", ">", "70000"), E.Comparison("?s", "<", "90000")]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("what,patterns,filters", QUERIES, ids=[q[0] for q in QUERIES])
def test_device_legacy_path_equals_the_restatement(ctx, what, patterns, filters):
    fx, tr = employee4_triples()
    # more data than the fixture: the same shape with 300 generated employees appended (non-integer and negative literals among them)
    rng = np.random.default_rng(4)
    for k in range(300):
        iri = f"http://example.org/gen{k}"
        sal = [str(int(rng.integers(30000, 150000))), f"{rng.integers(30000, 150000)}.5", "n/a", "-12"][k % 4] if k % 7 == 0 else str(int(rng.integers(30000, 150000)))
        tr += [(iri, FOAF + "name", iri), (iri, FOAF + "title", ["Developer", "Manager", "Salesperson"][k % 3]), (iri, FOAF + "workplaceHomepage", fx["workplace"]),
               (iri, DS + "annual_salary", sal)]
    db = E.SparqlDatabase(ctx=ctx)
    for t in tr:
        db.add_triple_parts(*t)
    ex = E.LegacyExecutor(db)
    want = L.execute_bgp(tr, patterns, filters)
    got = ex.execute_bgp(patterns, filters)
    keys = sorted({t for pt in patterns for t in (pt[0], pt[2]) if t.startswith("?")})
    assert bag(got, keys) == bag(want, keys), what
    for indexed in (True,):
        db.build_all_indexes()
        assert bag(ex.execute_bgp(patterns, filters), keys) == bag(want, keys), what + " (indexed)"


@pytest.mark.gpu
def test_device_legacy_path_rejects_what_it_does_not_evaluate(ctx):
    fx, tr = employee4_triples()
    db = E.SparqlDatabase(ctx=ctx)
    for t in tr:
        db.add_triple_parts(*t)
    ex = E.LegacyExecutor(db)
    with pytest.raises(c.KolibrieError) as e:  # '-' sends the operand to the reference's arithmetic parser
        ex.execute_bgp([("?e", DS + "annual_salary", "?s")], [E.Comparison("?s", ">", "-5")])
    assert e.value.status == c.KB_E_UNSUPPORTED
    with pytest.raises(c.KolibrieError):  # an IRI contains '/'
        ex.execute_bgp([("?e", FOAF + "name", "?n")], [E.Comparison("?n", "=", "http://example.org/employee1")])
