"""CPU: the vectorised generators assign ids exactly like a sequential Dictionary::encode pass (dictionary.rs:32-48)."""
import numpy as np

from kolibrie_b200 import datagen
from kolibrie_b200.engine import Dictionary


def sequential_employee(E, seed=42):
    d = Dictionary()
    rows = []
    for i in range(E):
        r1 = int(datagen.splitmix64_at(seed, np.array([2 * i], dtype=np.uint64))[0])
        r2 = int(datagen.splitmix64_at(seed, np.array([2 * i + 1], dtype=np.uint64))[0])
        iri = f"http://example.org/employee{i + 1}"
        title = datagen.POSITIONS[r1 % 3]
        sal = str(30000 + r2 % 120000)
        for p, o in (("foaf:name", iri), ("foaf:title", title), ("foaf:workplaceHomepage", "http://example.org/company"),
                     ("ds:full_or_part_time", "F"), ("ds:salary_or_hourly", "SALARY"), ("ds:annual_salary", sal)):
            rows.append((d.encode(iri), d.encode(p), d.encode(o)))
    return d, np.array(rows, dtype=np.uint32)


def test_employee_ids_match_sequential_encoder():
    for E in (1, 2, 7, 300):
        d, rows = sequential_employee(E)
        g = datagen.employee_dataset(E)
        assert np.array_equal(g.s, rows[:, 0]) and np.array_equal(g.p, rows[:, 1]) and np.array_equal(g.o, rows[:, 2])
        assert g.n_ids == len(d.id_to_string)
        num, isn = d.numeric_table()
        assert np.array_equal(num, g.num_or0) and np.array_equal(isn, g.is_num)
        for name in ("foaf:name", "foaf:title", "ds:annual_salary", "F", "SALARY"):
            assert g.ids[name] == d.lookup(name)
        assert g.ids["company"] == d.lookup("http://example.org/company")


def test_store_is_in_btreeset_order():
    g = datagen.employee_dataset(1000)
    key = (g.s.astype(np.uint64) << 40) | (g.p.astype(np.uint64) << 32) | g.o.astype(np.uint64)
    assert (np.diff(key.astype(np.int64)) > 0).all(), "document order == (s,p,o) order, the reference's BTreeSet iteration order"


def test_taxonomy_shape():
    t = datagen.taxonomy_dataset(fanout=3, depth=3, n_instances=100)
    assert t.n_classes == 40 and len(t.s) == 39 + 100
    sc = t.p == t.ids["rdfs:subClassOf"]
    assert sc.sum() == 39
    assert (t.o[sc] < t.s[sc]).all()  # parent ids precede children (breadth-first)


def test_employee_shard_closed_form_tail():
    """beyond the point where every salary literal has been seen, shard rows come from the closed form; they must equal the
    rows of the full sequential dataset"""
    E = 2_400_000
    full = datagen.employee_dataset(E)
    assert (full.sal_id_by_value >= 0).all()
    parts = [datagen.employee_shard(E, r, 2, prefix=2_000_000) for r in range(2)]
    from kolibrie_b200.dist import shard_of

    for r, part in enumerate(parts):
        keep = shard_of(full.s, 2) == r
        assert np.array_equal(part.s, full.s[keep]) and np.array_equal(part.p, full.p[keep]) and np.array_equal(part.o, full.o[keep])
    assert sum(p.n_employees for p in parts) == E


def test_closed_form_helpers_agree_with_the_sequential_dataset():
    """the closed forms bench.py asserts its multi-GPU legs against (subject id / title id of a global employee number, the employees
    of a shard, the reports_to companion relation) against the sequentially generated dataset"""
    E = 2_300_000
    full = datagen.employee_dataset(E)
    title_p = full.ids["foaf:title"]
    subj_of = full.s[full.p == title_p]      # subject of employee k, in generation order
    title_of = full.o[full.p == title_p]
    for world in (1, 2, 4):
        seen = []
        for r in range(world):
            part = datagen.employee_shard(E, r, world, prefix=2_000_000)
            idx = datagen.employee_indices_of_shard(part, r, world)
            seen.append(idx)
            assert np.array_equal(datagen.employee_subject_ids(part, idx), subj_of[idx])
            assert np.array_equal(datagen.employee_title_ids(part, idx), title_of[idx])
            assert np.array_equal(np.unique(part.s), np.sort(subj_of[idx]))            # exactly the shard's employees
            e, m, t = datagen.reports_to_relation(part, r, world)
            assert np.array_equal(e, subj_of[idx]) and len(m) == len(e)
            pos = {int(s): k for k, s in enumerate(subj_of[:50_000])}                  # spot check: m's title from the full dataset
            for k in np.nonzero(np.isin(m, subj_of[:50_000]))[0][:200]:
                assert int(title_of[pos[int(m[k])]]) == int(t[k])
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(E))


def test_permuted_and_multivalued_datasets():
    d = datagen.employee_dataset(3000)
    s, p, o, num, isn, pi = datagen.permuted_dataset(d)
    assert sorted(pi.tolist()) == list(range(d.n_ids))                                 # a permutation of all ids
    inv = np.empty_like(pi); inv[pi] = np.arange(d.n_ids, dtype=np.uint32)
    back = datagen.canonical_rows(np.stack([inv[s], inv[p], inv[o]], axis=1))
    assert np.array_equal(back, datagen.canonical_rows(np.stack([d.s, d.p, d.o], axis=1)))  # same triples, relabelled and shuffled
    assert np.array_equal(num[pi], d.num_or0) and np.array_equal(isn[pi], d.is_num)
    s, p, o, num, isn, meta = datagen.multivalued_dataset(5000, per_subject=3)
    tag = p == 1
    assert tag.sum() == 15000 and (p == 2).sum() == 5000 and (p == 3).sum() == 5000
    pairs = np.unique(np.stack([s[tag], o[tag]], axis=1), axis=0)
    assert len(pairs) == 15000                                                          # three DISTINCT tags per subject
    assert np.array_equal(num[o[p == 2]], meta["score"].astype(np.float64)) and isn[o[p == 2]].all()
