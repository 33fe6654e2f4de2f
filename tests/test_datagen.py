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
