"""GPU (-m gpu): id -> string decode of result columns on the device (kb_dict_strings_load / kb_rel_decode) against
Dictionary::decode (shared/src/dictionary.rs:50-52) and the final step of ExecutionEngine::execute (engine.rs:27-51)."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200.engine import Constant, ExecutionEngine, SparqlDatabase, StarJoin, TableScan, Variable

pytestmark = pytest.mark.gpu


def make_dictionary(rng, n):
    alphabet = list("abcdefghijklmnopqrstuvwxyz0123456789:/#.-_") + ["é", "ß", "中", "🙂"]
    out = []
    for i in range(n):
        ln = int(rng.choice([0, 1, 2, 7, 8, 31, 32, 33, 64, 100, 300])) if i % 7 == 0 else int(rng.integers(0, 60))
        out.append("".join(rng.choice(alphabet, ln)) if ln else "")
    return out


@pytest.mark.parametrize("n_rows", [0, 1, 31, 32, 33, 5000, 70001])
def test_decode_column_vs_dictionary(ctx, n_rows):
    rng = np.random.default_rng(n_rows + 1)
    strings = make_dictionary(rng, 3000)
    ctx.dict_strings_load(strings)
    ids = rng.integers(0, 3000, n_rows).astype(np.uint32)
    if n_rows > 10:
        ids[::9] = rng.integers(3000, 1 << 30, len(ids[::9]))  # ids the dictionary does not hold -> "unknown" (engine.rs:44)
    other = rng.integers(0, 3000, n_rows).astype(np.uint32)
    rel = ctx.rel_from_host([4, 9], [ids, other])
    for col, src in ((0, ids), (1, other)):
        off, data = rel.decode(col)
        want = [(strings[i] if i < len(strings) else "unknown").encode("utf-8") for i in src.tolist()]
        assert off[0] == 0 and len(off) == n_rows + 1
        assert np.array_equal(np.diff(off.astype(np.int64)), np.array([len(w) for w in want], dtype=np.int64))
        assert data.tobytes() == b"".join(want)
    assert rel.decode_strings(0) == [(strings[i] if i < len(strings) else "unknown") for i in ids.tolist()]


def test_decode_without_dictionary_and_quoted_ids(ctx):
    ctx.dict_strings_load([])
    rel = ctx.rel_from_host([0], [np.array([0, 5, 6], np.uint32)])
    assert rel.decode_strings(0) == ["unknown"] * 3
    ctx.dict_strings_load(["a", "bc"])
    quoted = ctx.rel_from_host([0], [np.array([1, 0x80000001], np.uint32)])
    with pytest.raises(c.KolibrieError) as e:
        quoted.decode(0)
    assert e.value.status == c.KB_E_UNSUPPORTED
    assert ctx.rel_from_host([0], [np.array([1, 0, 1], np.uint32)]).decode_strings(0) == ["bc", "a", "bc"]


def test_execute_returns_strings_like_the_reference(ctx):
    """ExecutionEngine::execute (engine.rs:27-51): ids all the way, strings only at the very end"""
    db = SparqlDatabase(ctx)
    people = [("http://example.org/employee%d" % i, ["Manager", "Developer", "Salesperson"][i % 3], str(30000 + 977 * i)) for i in range(1, 41)]
    for iri, title, sal in people:
        db.add_triple_parts(iri, "foaf:name", iri)
        db.add_triple_parts(iri, "foaf:title", title)
        db.add_triple_parts(iri, "ds:annual_salary", sal)
    d = db.dictionary
    e, t, s = Variable("?e"), Variable("t"), Variable("s")
    op = StarJoin("e", [(e, Constant(d.lookup("foaf:title")), t), (e, Constant(d.lookup("ds:annual_salary")), s)])
    rows = ExecutionEngine.execute(op, db)
    want = sorted((iri, title, sal) for iri, title, sal in people)
    assert sorted((r["e"], r["t"], r["s"]) for r in rows) == want
    ids = ExecutionEngine.execute_with_ids(op, db)
    assert sorted((d.decode(r["e"]), d.decode(r["t"]), d.decode(r["s"])) for r in ids) == want
    one = ExecutionEngine.execute(TableScan((e, Constant(d.lookup("foaf:title")), Constant(d.lookup("Manager")))), db)
    assert sorted(r["e"] for r in one) == sorted(iri for iri, title, _ in people if title == "Manager")
