"""Known answers the reference's own tests hold for RDF-star scans, DELETE WHERE and the per-firing window evaluation of the RSP engine
(tests/golden/rdf_star.json, rsp_windows.json — sources cited there), replayed through the host mirror of the reference's interfaces
(kolibrie_b200/engine.py): on the oracle in the CPU suite, on the device with -m gpu (index on and off). Plus the QuotedTripleStore unit
tests of shared/src/quoted_triple_store.rs:82-157 against its mirror."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import engine as E
from tests import helpers as H
from tests.oracle_ctx import OracleCtx

STAR = H.load("rdf_star.json")
RSP = H.load("rsp_windows.json")


def backends():
    return [pytest.param("oracle", id="oracle"), pytest.param("device", marks=pytest.mark.gpu, id="device"),
            pytest.param("device+index", marks=pytest.mark.gpu, id="device+index")]


@pytest.fixture
def make_db(request):
    made = []

    def _make(kind):
        if kind == "oracle":
            db = E.SparqlDatabase(ctx=OracleCtx())
        else:
            db = E.SparqlDatabase(ctx=request.getfixturevalue("ctx"))
        db._kind = kind
        made.append(db)
        return db

    yield _make


def sync(db):
    db._sync()
    if db._kind == "device+index":
        db.ctx.build_index()


def data_term(t):
    return tuple(data_term(x) for x in t) if isinstance(t, list) else t


def pattern_term(db, t):
    if isinstance(t, list):
        return E.QuotedTriple(*(pattern_term(db, x) for x in t))
    if t.startswith("?"):
        return E.Variable(t)
    return E.Constant(db.dictionary.encode(t))


def run_scan(db, pattern, select):
    sync(db)
    op = E.Projection(E.TableScan(tuple(pattern_term(db, t) for t in pattern)), ["?" + v for v in select])
    db._uploaded_version = db._version  # encode() of query constants does not change the triples
    rows = E.ExecutionEngine.execute(op, db)
    return sorted(tuple(r[v] for v in select) for r in rows)


@pytest.mark.parametrize("kind", backends())
@pytest.mark.parametrize("case", STAR["cases"], ids=[x["name"] for x in STAR["cases"]])
def test_rdf_star_scans(make_db, kind, case):
    db = make_db(kind)
    for s, p, o in case["data"]:
        db.add_statement(data_term(s), data_term(p), data_term(o))
    got = run_scan(db, case["pattern"], case["select"])
    assert got == sorted(tuple(r) for r in case["rows"])
    if "n_rows" in case:
        assert len(got) == case["n_rows"]
    if "subject_of_t" in case:  # rdf_star_test.rs:281-329: SUBJECT(?t) of the bound quoted triple is alice
        ids = E.ExecutionEngine.execute_with_ids(E.TableScan(tuple(pattern_term(db, t) for t in case["pattern"])), db)
        assert len(ids) == 1 and E.is_quoted_triple_id(ids[0]["t"])
        comp = db.quoted_triple_store.decode(ids[0]["t"])
        assert db.dictionary.decode(comp[0]) == case["subject_of_t"]
        # FILTER(isTRIPLE(?t)) keeps the row (types.rs:170-183)
        f = E.Filter(E.TableScan(tuple(pattern_term(db, t) for t in case["pattern"])), E.Condition(E.FunctionCall("isTRIPLE", ["?t"])))
        assert len(E.ExecutionEngine.execute_with_ids(f, db)) == 1


@pytest.mark.parametrize("kind", backends())
def test_quoted_scan_semantics_beyond_the_fixtures(make_db, kind):
    """resolve_quoted_triple_scan (engine.rs:1111-1188) corner cases: a variable shared between the quoted term and the outer pattern is
    a join condition (conflicting rows are dropped), a repeated variable inside the quoted term must bind consistently, a nested
    quoted term matches any quoted-triple id, a quoted term in object position"""
    db = make_db(kind)
    ex = "http://example.org/"
    db.add_statement((ex + "a", ex + "says", ex + "a"), ex + "by", ex + "a")       # << a says a >> by a
    db.add_statement((ex + "a", ex + "says", ex + "b"), ex + "by", ex + "b")       # << a says b >> by b
    db.add_statement((ex + "b", ex + "says", ex + "a"), ex + "by", ex + "c")       # << b says a >> by c
    db.add_statement(ex + "d", ex + "cites", (ex + "a", ex + "says", ex + "b"))    # d cites << a says b >>
    db.add_statement(((ex + "a", ex + "says", ex + "b"), ex + "in", ex + "g"), ex + "by", ex + "e")  # << << a says b >> in g >> by e
    says, by = ex + "says", ex + "by"
    assert run_scan(db, [["?x", says, "?y"], by, "?y"], ["x", "y"]) == [(ex + "a", ex + "a"), (ex + "a", ex + "b")]   # outer object joins with inner ?y
    assert run_scan(db, [["?x", says, "?x"], by, "?w"], ["x", "w"]) == [(ex + "a", ex + "a")]                         # repeated inner variable
    assert run_scan(db, ["?who", ex + "cites", ["?x", says, "?y"]], ["who", "x", "y"]) == [(ex + "d", ex + "a", ex + "b")]
    assert run_scan(db, [[["?p", "?q", "?r"], ex + "in", "?g"], by, "?w"], ["g", "w"]) == [(ex + "g", ex + "e")]       # nested term: any quoted id
    assert run_scan(db, [["?x", ex + "nothing", "?y"], by, "?w"], ["x"]) == []


@pytest.mark.parametrize("kind", backends())
def test_delete_where(make_db, kind):
    """rdf_star_test.rs:384-405: DELETE WHERE { ?s knows ?o } — the WHERE scan on the hot path, then delete_triple per match"""
    fx = STAR["delete_where"]
    db = make_db(kind)
    for s, p, o in fx["data"]:
        db.add_statement(s, p, o)
    assert len(db.triples) == fx["triples_before"]
    pat = tuple(pattern_term(db, t) for t in fx["delete_pattern"])
    sync(db)
    rows = E.ExecutionEngine.execute_with_ids(E.TableScan(pat), db)
    pid = pat[1].id
    for r in rows:
        assert db.delete_triple((r["s"], pid, r["o"]))
    assert len(db.triples) == fx["triples_after"]
    sync(db)
    assert E.ExecutionEngine.execute_with_ids(E.TableScan(pat), db) == []
    if kind != "oracle":  # the device store itself: kb_store_delete by value leaves the same single triple
        ctx = db.ctx
        d2 = E.Dictionary()
        tr = np.array([[d2.encode(x) for x in t] for t in fx["data"]], dtype=np.uint32)
        ctx.store_load(tr[:, 0], tr[:, 1], tr[:, 2])
        knows = tr[tr[:, 1] == d2.lookup(fx["delete_pattern"][1])]
        ctx.store_delete(knows[:, 0], knows[:, 1], knows[:, 2])
        assert ctx.store_size()[0] == fx["triples_after"]


@pytest.mark.parametrize("kind", backends())
@pytest.mark.parametrize("case", RSP["cases"], ids=[x["name"] for x in RSP["cases"]])
def test_rsp_window_firings(make_db, kind, case):
    """per firing: evict the previous window's triples, add the current window's (rsp_engine.rs:94-104), evaluate the query; the R2S
    operator (out of the hot path) turns the per-firing results into the emitted rows: ISTREAM = rows not in the previous firing's
    result, RSTREAM = all rows"""
    db = make_db(kind)
    last_rows, prev_window = set(), []
    for f in case["firings"]:
        for t in prev_window:
            db.delete_triple(tuple(db.dictionary.encode(x) for x in t))
        for t in f["window"]:
            db.add_statement(*t)
        prev_window = f["window"]
        rows = set(run_scan(db, case["pattern"], case["select"]))
        assert len(rows) == len(f["window"])
        emit = rows - last_rows if case["stream"] == "ISTREAM" else rows
        assert sorted(emit) == sorted(tuple(r) for r in f["emit"]), (case["name"], f)
        last_rows = rows


@pytest.mark.gpu
@pytest.mark.parametrize("case", RSP["cases"], ids=[x["name"] for x in RSP["cases"]])
def test_rsp_window_firings_as_device_segments(ctx, case):
    """the same firings through the device store's own window maintenance: kb_store_evict(previous tag) + kb_store_append(tag)"""
    d = E.Dictionary()
    pat = []
    slots = E.SlotMap()
    for t in case["pattern"]:
        pat.append(c.V(slots.of(t)) if t.startswith("?") else c.K(d.encode(t)))
    ctx.store_clear()
    last, tag = set(), None
    for i, f in enumerate(case["firings"]):
        tr = np.array([[d.encode(x) for x in t] for t in f["window"]], dtype=np.uint32)
        if tag is not None:
            ctx.store_evict(tag)
        tag = 100 + i
        ctx.store_append(tr[:, 0], tr[:, 1], tr[:, 2], tag)
        rel = ctx.scan([c.pattern(*pat)])[0]
        order = [slots.of("?" + v) for v in case["select"]]
        rows = {tuple(d.decode(int(x)) for x in r) for r in rel.to_numpy(order)}
        emit = rows - last if case["stream"] == "ISTREAM" else rows
        assert sorted(emit) == sorted(tuple(r) for r in f["emit"])
        last = rows


def test_quoted_triple_store_unit_tests():
    """shared/src/quoted_triple_store.rs:82-157, test for test"""
    S = E.QuotedTripleStore
    st = S()
    i = st.encode(1, 2, 3)
    assert E.is_quoted_triple_id(i) and st.decode(i) == (1, 2, 3)                      # test_encode_decode_roundtrip
    st = S()
    assert st.encode(1, 2, 3) == st.encode(1, 2, 3) and len(st) == 1                    # test_deduplication
    st = S()
    assert st.encode(1, 2, 3) != st.encode(4, 5, 6) and len(st) == 2                    # test_different_triples_get_different_ids
    st = S()
    inner = st.encode(1, 2, 3)
    outer = st.encode(inner, 4, 5)                                                      # test_nested_quoted_triples
    assert E.is_quoted_triple_id(inner) and E.is_quoted_triple_id(outer) and inner != outer and st.decode(outer) == (inner, 4, 5)
    for v, want in ((0, False), (100, False), (0x7FFF_FFFF, False), (0x8000_0000, True), (0x8000_0001, True), (0xFFFF_FFFF, True)):
        assert E.is_quoted_triple_id(v) is want                                         # test_is_quoted_triple_id
    s1 = S()
    id1 = s1.encode(1, 2, 3)
    s2 = S()
    s2.next_qt_id = s1.next_qt_id
    id2 = s2.encode(4, 5, 6)
    s1.merge(s2)                                                                        # test_merge
    assert len(s1) == 2 and s1.decode(id1) == (1, 2, 3) and s1.decode(id2) == (4, 5, 6)
    assert S().decode(0x8000_0000) is None                                              # test_decode_nonexistent
    assert S().is_empty() and S().next_qt_id == 0x8000_0000
