"""Seeded synthetic inputs in the reference's shapes (SURVEY.md §8d / BASELINE.md §4).

The reference's generator (kolibrie/examples/synthetic_data/gen_data.rs:100-150) draws from an OS-seeded RNG and its
datasets are Git-LFS pointers, so the shapes are regenerated here deterministically:

* employee shape — 6 triples per employee in document order (gen_data.rs:116-143): foaf:name (= the employee IRI string, so
  it gets the SAME dictionary id as the subject), foaf:title in {Manager, Developer, Salesperson}, foaf:workplaceHomepage
  (one company), ds:full_or_part_time "F", ds:salary_or_hourly "SALARY", ds:annual_salary in [30000, 150000).
  PRNG = splitmix64, seed 42, two draws per employee (title, salary). Dictionary ids are assigned in first-seen order,
  subject, predicate, object per triple (kolibrie/src/sparql_database.rs:668-673; shared/src/dictionary.rs:32-48).
* taxonomy shape (config 4) — complete `fanout`-ary class tree + rdf:type facts on random classes, seed 43.

Everything is vectorised numpy so the 100 M-triple shape generates in seconds; ids come out exactly as a sequential
Dictionary::encode pass would assign them (checked against a literal sequential encoder in tests/test_datagen.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

POSITIONS = ("Manager", "Developer", "Salesperson")  # gen_data.rs:22-24
_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_at(seed: int, k: np.ndarray) -> np.ndarray:
    """k-th output (0-based) of a splitmix64 stream seeded with `seed` (state advances by GAMMA per draw)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (k.astype(np.uint64) + np.uint64(1)) * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


@dataclass
class EmployeeData:
    s: np.ndarray
    p: np.ndarray
    o: np.ndarray
    n_ids: int
    ids: Dict[str, int]          # named constants: predicates, titles, company, "F", "SALARY"
    num_or0: np.ndarray          # float64[n_ids]  parse::<f64>().unwrap_or(0.0) of every term
    is_num: np.ndarray           # uint8[n_ids]
    n_employees: int
    salary_of_employee: np.ndarray = field(repr=False, default=None)
    title_of_employee: np.ndarray = field(repr=False, default=None)
    sal_id_by_value: np.ndarray = field(repr=False, default=None)    # id of the literal str(30000+k), -1 if never drawn
    title_id_by_value: np.ndarray = field(repr=False, default=None)  # id of POSITIONS[k], -1 if never drawn
    # closed form of the GLOBAL dataset this shard was cut from (employee_subject_ids): subject ids of its first `len(subj_prefix)`
    # employees; every later employee i has id prefix_ids + (i - len(subj_prefix))
    subj_prefix: np.ndarray = field(repr=False, default=None)
    prefix_ids: int = 0
    n_total: int = 0
    seed: int = 42

    @property
    def n_triples(self) -> int:
        return len(self.s)


def employee_dataset(n_employees: int, seed: int = 42, first: int = 1, global_ids: bool = True) -> EmployeeData:
    """Employees first..first+n-1 of the seeded stream. With first=1 the ids are those of a fresh dictionary."""
    E = int(n_employees)
    assert E >= 1 and first == 1, "sub-ranges are produced by employee_shard()"
    i = np.arange(E, dtype=np.uint64)
    r_title = splitmix64_at(seed, 2 * i)
    r_sal = splitmix64_at(seed, 2 * i + 1)
    title_idx = (r_title % np.uint64(3)).astype(np.int64)
    salary = (np.uint64(30000) + r_sal % np.uint64(120000)).astype(np.int64)

    # first-seen flags
    new_title = np.zeros(E, dtype=bool)
    _, first_t = np.unique(title_idx, return_index=True)
    new_title[first_t] = True
    new_sal = np.zeros(E, dtype=bool)
    uniq_sal, first_s = np.unique(salary, return_index=True)
    new_sal[first_s] = True

    # ids consumed by each employee: employee 0 also introduces 6 predicates, the company, "F" and "SALARY"
    consumed = 1 + new_title.astype(np.int64) + new_sal.astype(np.int64)
    consumed[0] += 9
    start = np.concatenate(([0], np.cumsum(consumed)[:-1]))  # id of each employee's subject
    n_ids = int(start[-1] + consumed[-1])
    subj = start.copy()

    # constants (document order inside employee 0: s, name, [name obj = s], title, TITLE, wh, company, fpt, F, soh, SALARY, sal, SALVAL)
    ids = {
        "foaf:name": 1, "foaf:title": 2, "foaf:workplaceHomepage": 4, "company": 5, "ds:full_or_part_time": 6, "F": 7,
        "ds:salary_or_hourly": 8, "SALARY": 9, "ds:annual_salary": 10,
    }
    # id of a new title introduced by employee e: right after the title predicate for e=0, right after the subject otherwise
    title_id_by_value = np.zeros(3, dtype=np.int64)
    for t in range(3):
        hits = np.nonzero(new_title & (title_idx == t))[0]
        if len(hits):
            e = int(hits[0])
            title_id_by_value[t] = 3 if e == 0 else subj[e] + 1
        else:
            title_id_by_value[t] = -1
    for t, name in enumerate(POSITIONS):
        if title_id_by_value[t] >= 0:
            ids[name] = int(title_id_by_value[t])
    # id of a new salary literal introduced by employee e
    sal_new_id = subj + 1 + new_title.astype(np.int64)
    sal_new_id[0] = 11
    sal_id_by_value = np.full(120000, -1, dtype=np.int64)
    sal_id_by_value[uniq_sal - 30000] = sal_new_id[first_s]

    title_obj = title_id_by_value[title_idx]
    sal_obj = sal_id_by_value[salary - 30000]

    s = np.repeat(subj, 6).astype(np.uint32)
    p = np.tile(np.array([1, 2, 4, 6, 8, 10], dtype=np.uint32), E)
    o = np.empty(6 * E, dtype=np.uint32)
    o[0::6] = subj
    o[1::6] = title_obj
    o[2::6] = 5
    o[3::6] = 7
    o[4::6] = 9
    o[5::6] = sal_obj

    num = np.zeros(n_ids, dtype=np.float64)
    isn = np.zeros(n_ids, dtype=np.uint8)
    num[sal_new_id[first_s]] = uniq_sal.astype(np.float64)
    isn[sal_new_id[first_s]] = 1
    return EmployeeData(s, p, o, n_ids, ids, num, isn, E, salary_of_employee=salary, title_of_employee=title_idx,
                        sal_id_by_value=sal_id_by_value, title_id_by_value=title_id_by_value,
                        subj_prefix=subj.astype(np.uint32), prefix_ids=n_ids, n_total=E, seed=seed)


SHARD_BLOCK_BITS = 10  # == KB_SHARD_BLOCK_BITS


def shard_of_np(keys: np.ndarray, world: int) -> np.ndarray:
    """kb_shard_of on an array: block-cyclic on the dense dictionary id, (key >> 10) % world"""
    return ((np.asarray(keys, dtype=np.uint32) >> np.uint32(SHARD_BLOCK_BITS)) % np.uint32(world)).astype(np.int64)


def mix32_np(x: np.ndarray) -> np.ndarray:
    """kb_shard_of's hash (murmur3 finaliser) on a uint32 array"""
    with np.errstate(over="ignore"):
        x = x.astype(np.uint32).copy()
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(13)
        x *= np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    return x


def employee_shard(n_total: int, rank: int, world: int, seed: int = 42, prefix: int = 4_000_000) -> EmployeeData:
    """The triples of the GLOBAL n_total-employee dataset whose subject id belongs to `rank` (kb_shard_of(subject, world)), in global
    document order — i.e. one GPU's shard under hash(subject) sharding. Ids are those of the global dictionary:
    once every title and salary literal has been seen (a few million employees) each further employee consumes exactly one id,
    so the tail is closed-form and a rank never materialises the other ranks' rows."""
    if world == 1:
        return employee_dataset(n_total, seed)
    P = min(prefix, n_total)
    while True:
        head = employee_dataset(P, seed)
        complete = (head.sal_id_by_value >= 0).all() and (head.title_id_by_value >= 0).all()
        if complete or P == n_total:
            break
        P = min(P * 2, n_total)
    subj_head = head.s[0::6]
    keep = shard_of_np(subj_head, world) == rank
    rows = np.repeat(keep, 6)
    S, Pp, Oo = [head.s[rows]], [head.p[rows]], [head.o[rows]]
    n_emp = int(keep.sum())
    chunk = 8_000_000
    preds = np.array([1, 2, 4, 6, 8, 10], dtype=np.uint32)
    for a in range(P, n_total, chunk):
        b = min(a + chunk, n_total)
        i = np.arange(a, b, dtype=np.uint64)
        subj = (np.uint64(head.n_ids) + (i - np.uint64(P))).astype(np.uint32)
        k = shard_of_np(subj, world) == rank
        i, subj = i[k], subj[k]
        r_title = splitmix64_at(seed, 2 * i)
        r_sal = splitmix64_at(seed, 2 * i + 1)
        t_obj = head.title_id_by_value[(r_title % np.uint64(3)).astype(np.int64)]
        s_obj = head.sal_id_by_value[(r_sal % np.uint64(120000)).astype(np.int64)]
        m = len(subj)
        o = np.empty(6 * m, dtype=np.uint32)
        o[0::6] = subj; o[1::6] = t_obj; o[2::6] = 5; o[3::6] = 7; o[4::6] = 9; o[5::6] = s_obj
        S.append(np.repeat(subj, 6)); Pp.append(np.tile(preds, m)); Oo.append(o)
        n_emp += m
    n_ids = head.n_ids + (n_total - P)
    return EmployeeData(np.concatenate(S), np.concatenate(Pp), np.concatenate(Oo), n_ids, head.ids, head.num_or0, head.is_num, n_emp,
                        sal_id_by_value=head.sal_id_by_value, title_id_by_value=head.title_id_by_value,
                        subj_prefix=subj_head.astype(np.uint32), prefix_ids=head.n_ids, n_total=n_total, seed=seed)


def employee_subject_ids(d: EmployeeData, index: np.ndarray) -> np.ndarray:
    """dictionary id of the subject of global employee number `index` (0-based) of the dataset `d` was cut from"""
    index = np.asarray(index, dtype=np.int64)
    P = len(d.subj_prefix)
    out = (np.int64(d.prefix_ids) + (index - P)).astype(np.uint32)
    head = index < P
    out[head] = d.subj_prefix[index[head]]
    return out


def employee_title_ids(d: EmployeeData, index: np.ndarray) -> np.ndarray:
    """dictionary id of the foaf:title object of global employee number `index` (closed form: first draw of the employee's pair)"""
    r = splitmix64_at(d.seed, 2 * np.asarray(index, dtype=np.uint64))
    return d.title_id_by_value[(r % np.uint64(3)).astype(np.int64)].astype(np.uint32)


def employee_indices_of_shard(d: EmployeeData, rank: int, world: int) -> np.ndarray:
    """global employee numbers whose subject belongs to shard `rank` (the employees of employee_shard(n_total, rank, world)), in order"""
    if world == 1:
        return np.arange(d.n_total, dtype=np.int64)
    P = len(d.subj_prefix)
    head = np.nonzero(shard_of_np(d.subj_prefix, world) == rank)[0].astype(np.int64)
    tail_i = np.arange(P, d.n_total, dtype=np.int64)
    tail_ids = (np.int64(d.prefix_ids) + (tail_i - P)).astype(np.uint32)
    return np.concatenate([head, tail_i[shard_of_np(tail_ids, world) == rank]])


def reports_to_relation(d: EmployeeData, rank: int, world: int, seed: int = 77):
    """A path-join companion of the employee shape (bench.py's shuffle leg): (?e ds:reports_to ?m) for the employees of this shard,
    ?m = a pseudo-random employee of the GLOBAL dataset — the rows live with ?e but join with (?m foaf:title ?t), which lives on the
    owner of ?m: a non-subject-key join. Returns (e ids, m ids, closed-form title id of m)."""
    idx = employee_indices_of_shard(d, rank, world)
    m_idx = (splitmix64_at(seed, idx.astype(np.uint64)) % np.uint64(d.n_total)).astype(np.int64)
    return employee_subject_ids(d, idx), employee_subject_ids(d, m_idx), employee_title_ids(d, m_idx)


def employee_queries(d: EmployeeData):
    """The BASELINE.md §4 queries on this dataset as (join_slot, patterns, filter) in capi terms. Slots: e=0 t=1 s=2 n=3 c=4."""
    from . import capi as c

    ids = d.ids
    e, t, s, n, co = 0, 1, 2, 3, 4
    q = {
        # cfg1: ?p foaf:workplaceHomepage ?c . ?p foaf:name ?n
        "cfg1": (e, [c.pattern(c.V(e), c.K(ids["foaf:workplaceHomepage"]), c.V(co)), c.pattern(c.V(e), c.K(ids["foaf:name"]), c.V(n))], []),
        # cfg2: ?e foaf:title ?t . ?e ds:annual_salary ?s . ?e foaf:name ?n FILTER(?s > 100000)
        "cfg2": (e, [c.pattern(c.V(e), c.K(ids["foaf:title"]), c.V(t)), c.pattern(c.V(e), c.K(ids["ds:annual_salary"]), c.V(s)),
                     c.pattern(c.V(e), c.K(ids["foaf:name"]), c.V(n))], [c.fop(c.F_CMP_NUM, slot=s, cmp=c.CMP_GT, value=100000.0)]),
        # cfg3: cfg2's patterns + ?e foaf:workplaceHomepage ?c (GROUP BY ?t COUNT is applied by the caller)
        "cfg3": (e, [c.pattern(c.V(e), c.K(ids["foaf:title"]), c.V(t)), c.pattern(c.V(e), c.K(ids["ds:annual_salary"]), c.V(s)),
                     c.pattern(c.V(e), c.K(ids["foaf:name"]), c.V(n)), c.pattern(c.V(e), c.K(ids["foaf:workplaceHomepage"]), c.V(co))], []),
    }
    q["star3"] = (e, q["cfg2"][1], [])  # the 3-pattern star without FILTER (SURVEY §8d worked numbers)
    return q


@dataclass
class TaxonomyData:
    s: np.ndarray
    p: np.ndarray
    o: np.ndarray
    n_ids: int
    ids: Dict[str, int]
    n_classes: int
    n_instances: int


def taxonomy_dataset(fanout: int = 10, depth: int = 6, n_instances: int = 48_888_890, seed: int = 43, first_instance: int = 0) -> TaxonomyData:
    """Complete `fanout`-ary class tree of `depth` levels below the root (subClassOf child->parent) + `x rdf:type C` facts.
    ids: rdfs:subClassOf = 0, rdf:type = 1, classes 2..2+n_classes-1 (breadth-first, root first), instances after."""
    n_classes = sum(fanout ** k for k in range(depth + 1))
    SC, TYPE, C0 = 0, 1, 2
    child = np.arange(1, n_classes, dtype=np.int64)
    parent = (child - 1) // fanout
    X0 = C0 + n_classes
    j = np.arange(first_instance, first_instance + n_instances, dtype=np.uint64)  # a rank's slice of the global instance range
    cls = (splitmix64_at(seed, j) % np.uint64(n_classes)).astype(np.int64)
    s = np.concatenate([child + C0, j.astype(np.int64) + X0]).astype(np.uint32)
    p = np.concatenate([np.full(len(child), SC), np.full(n_instances, TYPE)]).astype(np.uint32)
    o = np.concatenate([parent + C0, cls + C0]).astype(np.uint32)
    return TaxonomyData(s, p, o, int(X0 + first_instance + n_instances), {"rdfs:subClassOf": SC, "rdf:type": TYPE}, n_classes, n_instances)


def taxonomy_rules(d: TaxonomyData):
    """R1 (?a sc ?b),(?b sc ?c)->(?a sc ?c); R2 (?x type ?a),(?a sc ?b)->(?x type ?b) (R2 = deep_taxonomy.rs:70-91)."""
    from . import capi as c

    sc, ty = d.ids["rdfs:subClassOf"], d.ids["rdf:type"]
    a, b, cc, x = 0, 1, 2, 3
    r1 = {"premise": [c.pattern(c.V(a), c.K(sc), c.V(b)), c.pattern(c.V(b), c.K(sc), c.V(cc))], "conclusion": [c.pattern(c.V(a), c.K(sc), c.V(cc))]}
    r2 = {"premise": [c.pattern(c.V(x), c.K(ty), c.V(a)), c.pattern(c.V(a), c.K(sc), c.V(b))], "conclusion": [c.pattern(c.V(x), c.K(ty), c.V(b))]}
    return [r1, r2]


def canonical_rows(a: np.ndarray) -> np.ndarray:
    """rows sorted lexicographically — the canonical form parity is judged on (the reference's row order is hash order)."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.shape[0] == 0 or a.shape[1] == 0:
        return a
    order = np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))
    return a[order]


def row_checksums(a: np.ndarray):
    """(row count, sum of mix64(row) mod 2^64, xor of mix64(row)) — order-independent digest for results too large to sort-compare."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    with np.errstate(over="ignore"):
        h = np.full(a.shape[0], 0x243F6A8885A308D3, dtype=np.uint64)
        for k in range(a.shape[1]):
            h = h ^ a[:, k].astype(np.uint64)
            h ^= h >> np.uint64(33)
            h *= np.uint64(0xFF51AFD7ED558CCD)
            h ^= h >> np.uint64(33)
            h *= np.uint64(0xC4CEB9FE1A85EC53)
            h ^= h >> np.uint64(33)
        return int(a.shape[0]), int(h.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(h)) if len(h) else 0


def permuted_dataset(d: EmployeeData, seed: int = 5):
    """The same dataset under a random relabelling of ALL dictionary ids and a random order of the triples: what a store looks like
    whose dictionary was not filled in document order (ids of one employee's terms far apart, subjects in no order). Returns
    (s, p, o, num_or0, is_num, pi) with pi[old id] = new id."""
    rng = np.random.default_rng(seed)
    pi = rng.permutation(d.n_ids).astype(np.uint32)
    order = rng.permutation(d.n_triples)
    num = np.zeros_like(d.num_or0)
    isn = np.zeros_like(d.is_num)
    num[pi] = d.num_or0
    isn[pi] = d.is_num
    return pi[d.s][order], pi[d.p][order], pi[d.o][order], num, isn, pi


def multivalued_dataset(n_subjects: int, per_subject: int = 3, seed: int = 9):
    """A store with one MULTI-VALUED predicate: subject k (id 100 + k) has `per_subject` distinct `tag` objects, one functional
    `score` literal and one functional `name`. ids: tag = 1, score = 2, name = 3; tag objects 10 .. 10 + 50; subjects from 100; score
    literals after the subjects (value v has id lit0 + v, v in [0, 1000)). Returns (s, p, o, num_or0, is_num, meta)."""
    n = int(n_subjects)
    subj = (np.arange(n, dtype=np.uint32) + np.uint32(100))
    k = np.arange(n, dtype=np.uint64)
    lit0 = 100 + n
    score = (splitmix64_at(seed, k) % np.uint64(1000)).astype(np.uint32)
    tag0 = (splitmix64_at(seed + 1, k) % np.uint64(50)).astype(np.uint32)
    assert 1 <= per_subject <= 7  # (tag0 + 7 j) mod 50 are distinct for j < 50 / 7
    tags = [tag0 + np.uint32(10)] + [((tag0 + np.uint32(j * 7)) % np.uint32(50)) + np.uint32(10) for j in range(1, per_subject)]
    s = np.concatenate([np.repeat(subj, per_subject), subj, subj])
    p = np.concatenate([np.full(n * per_subject, 1, np.uint32), np.full(n, 2, np.uint32), np.full(n, 3, np.uint32)])
    o = np.concatenate([np.stack(tags, axis=1).reshape(-1), score + np.uint32(lit0), subj])
    n_ids = lit0 + 1000
    num = np.zeros(n_ids, dtype=np.float64)
    isn = np.zeros(n_ids, dtype=np.uint8)
    num[lit0:] = np.arange(1000, dtype=np.float64)
    isn[lit0:] = 1
    return s.astype(np.uint32), p, o.astype(np.uint32), num, isn, {"subj": subj, "score": score, "tags": tags, "lit0": lit0, "per_subject": per_subject}
