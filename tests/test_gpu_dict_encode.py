"""GPU (-m gpu): kb_dict_encode — Dictionary::encode (shared/src/dictionary.rs:32-48) for a batch of terms on the device — against the
sequential encode of the host mirror: ids bit-exact (existing strings keep theirs, new strings get the next ids in first-seen order),
the appended strings decodable on the device (kb_rel_decode), over several batches, with duplicates, empty / long / non-ASCII terms and
strings that share a prefix or a length."""
import numpy as np
import pytest

from kolibrie_b200 import capi as c
from kolibrie_b200 import engine as E

pytestmark = pytest.mark.gpu


def seq_encode(d: E.Dictionary, terms):
    return np.array([d.encode(t) for t in terms], dtype=np.uint32)


def batch(rng, n, vocab, tag):
    out = []
    for _ in range(n):
        k = int(rng.integers(0, vocab))
        kind = k % 5
        if kind == 0:
            out.append(f"http://example.org/{tag}/employee{k}")
        elif kind == 1:
            out.append(str(k * 37 % 1000))          # short numeric literals, many duplicates
        elif kind == 2:
            out.append("é" * (k % 7) + f"ünï{k}")   # multi-byte UTF-8
        elif kind == 3:
            out.append("x" * (k % 300))              # same bytes, different lengths (the empty string included)
        else:
            out.append(f"http://example.org/{tag}/employee{k}/name")  # shares a prefix with kind 0
    return out


def test_batches_match_sequential_encode(ctx):
    rng = np.random.default_rng(3)
    want_d = E.Dictionary()
    ctx.dict_strings_load([])
    total_new = 0
    for b in range(5):
        terms = batch(rng, 20000, 3000 * (b + 1), tag="a" if b < 3 else "b")
        before = len(want_d.id_to_string)
        want = seq_encode(want_d, terms)
        got, first = ctx.dict_encode(terms)
        assert np.array_equal(got, want), f"batch {b}: ids differ from the sequential encode"
        n_new = len(want_d.id_to_string) - before
        assert len(first) == n_new
        assert [terms[int(p)] for p in first] == want_d.id_to_string[before:], "first-seen positions name the new strings in id order"
        total_new += n_new
        n_ids, n_bytes = ctx.dict_strings_info()
        assert n_ids == len(want_d.id_to_string) and n_bytes == sum(len(s.encode()) for s in want_d.id_to_string)
    # the device dictionary decodes every id it handed out
    ids = np.arange(len(want_d.id_to_string), dtype=np.uint32)
    rel = ctx.rel_from_host([0], [ids])
    assert rel.decode_strings(0) == want_d.id_to_string
    # a batch of known terms only: no new ids, nothing appended
    again = [want_d.id_to_string[int(i)] for i in rng.integers(0, len(ids), 5000)]
    got, first = ctx.dict_encode(again)
    assert np.array_equal(got, seq_encode(want_d, again)) and len(first) == 0


def test_encode_extends_a_loaded_dictionary(ctx):
    d = E.Dictionary()
    for t in ("http://example.org/p", "Alice", "", "42", "Bob"):
        d.encode(t)
    ctx.dict_strings_load(d.id_to_string)
    terms = ["Bob", "Carol", "", "Carol", "42", "http://example.org/q", "Alice", "http://example.org/p", "Dave", "Carol"]
    got, first = ctx.dict_encode(terms)
    assert np.array_equal(got, seq_encode(d, terms))
    assert [int(p) for p in first] == [1, 5, 8]
    assert ctx.dict_strings_info()[0] == len(d.id_to_string) == 8


def test_bulk_load_through_the_host_mirror(ctx):
    """SparqlDatabase.add_triples_bulk (device encode) builds the same dictionary and triples as add_triple_parts one by one"""
    rng = np.random.default_rng(9)
    st = [(f"http://e.org/s{int(rng.integers(0, 400))}", f"http://e.org/p{int(rng.integers(0, 5))}", str(int(rng.integers(0, 90)))) for _ in range(3000)]
    a = E.SparqlDatabase(ctx=ctx)
    ctx.dict_strings_load([])
    a.add_triples_bulk(st[:1500])
    a.add_triples_bulk(st[1500:])
    b = E.SparqlDatabase(ctx=ctx)
    for s, p, o in st:
        b.add_triple_parts(s, p, o)
    assert a.dictionary.id_to_string == b.dictionary.id_to_string and a.triples == b.triples


def test_one_large_batch(ctx):
    """3 M terms of the employee shape (6 triples per employee, subjects repeated): ids equal the sequential encode"""
    n_emp = 170_000
    terms = []
    for e in range(n_emp):
        s = f"http://example.org/employee{e}"
        for p, o in (("foaf:name", s), ("foaf:title", "Developer" if e % 3 else "Manager"), ("ds:annual_salary", str(50000 + e * 7919 % 120000))):
            terms += [s, p, o]
    d = E.Dictionary()
    want = seq_encode(d, terms)
    ctx.dict_strings_load([])
    got, first = ctx.dict_encode(terms)
    assert np.array_equal(got, want) and len(first) == len(d.id_to_string)
