"""TEST INFRASTRUCTURE: the oracle behind the method names of kolibrie_b200.capi.Context, so that the HOST-side logic of
kolibrie_b200/engine.py (plan walking, FILTER compilation, quoted-triple resolution, slot mapping) runs in the CPU suite without a GPU.
The product never imports this: a SparqlDatabase is handed an OracleCtx explicitly by a test."""
from typing import List, Sequence

import numpy as np

from tests import oracle_api as O


class OracleRelation:
    def __init__(self, rel, strings):
        self.rel = rel
        self._strings = strings

    def info(self):
        return self.rel.n_rows, self.rel.slots

    @property
    def n_rows(self):
        return self.rel.n_rows

    @property
    def slots(self):
        return self.rel.slots

    def column(self, col: int) -> np.ndarray:
        return self.rel.to_numpy([self.rel.slots[col]])[:, 0].copy()

    def to_numpy(self, slot_order=None):
        return self.rel.to_numpy(slot_order)

    def decode_strings(self, col: int) -> List[str]:
        ids = self.column(col)
        if (ids & 0x80000000).any():
            from kolibrie_b200 import capi as c
            raise c.KolibrieError(c.KB_E_UNSUPPORTED, "the column holds quoted-triple ids (bit 31): decode them on the host")
        return [self._strings[i] if i < len(self._strings) else "unknown" for i in ids]  # engine.rs:44

    def free(self):
        self.rel = None


class OracleCtx:
    def __init__(self):
        self.db = None
        self._strings: List[str] = []
        self._num = None

    def store_load(self, s, p, o):
        self._spo = (np.asarray(s, np.uint32), np.asarray(p, np.uint32), np.asarray(o, np.uint32))
        self._make()

    def dict_numeric_load(self, num, isn):
        self._num = (np.asarray(num, np.float64), np.asarray(isn, np.uint8))
        self._make()

    def _make(self):
        if getattr(self, "_spo", None) is None:
            return
        self.db = O.Db(*self._spo, *(self._num if self._num is not None else (None, None)))

    def dict_strings_load(self, strings: Sequence[str]):
        self._strings = list(strings)

    def build_index(self):
        return 0, 0.0

    def _wrap(self, rel):
        return OracleRelation(rel, self._strings)

    def scan(self, pats, pushdown=None):
        return [self._wrap(self.db.scan(p, pushdown[i] if pushdown and i < len(pushdown) else None)) for i, p in enumerate(pats)]

    def filter(self, rel, ops):
        return self._wrap(self.db.filter(rel.rel, ops))

    def project(self, rel, slots):
        return self._wrap(O.project(rel.rel, list(slots)))

    def hash_join(self, left, right):
        return self._wrap(O.hash_join(left.rel, right.rel))

    def bind_join(self, left, pat):
        return self._wrap(O.hash_join(left.rel, self.db.scan(pat)))

    def star_join(self, join_slot, pats, filt=None):
        return self._wrap(self.db.bgp(pats, filt))

    def rel_from_host(self, slots, cols):
        return self._wrap(O.rel_from_host(list(slots), [np.asarray(x, np.uint32) for x in cols]))

    # ---- Datalog (Reasoner mirror on CPU): the contracts of kb_datalog_fixpoint / kb_datalog_fixpoint_seed restated with the oracle
    class _Stats:
        def __init__(self, w=None):
            self.rounds = len(w["round_new"]) if w else 0
            self.round_new = list(w["round_new"]) + [0] * 64 if w else [0] * 64
            self.inferred = len(w["facts"]) if w else 0
            self.derivations = int(w["derivations"]) if w else 0
            self.device_ms = 0.0

    def store_append(self, s, p, o, tag):
        self._spo = tuple(np.concatenate([a, np.asarray(b, np.uint32)]) for a, b in zip(self._spo, (s, p, o)))
        self._make()

    def _rows(self):
        return np.stack(self._spo, axis=1)

    def datalog_fixpoint(self, rules, strategy=0):
        from kolibrie_b200 import capi as c

        w = self.db.fixpoint(rules, strategy)
        if w["status"] != 0:
            raise c.KolibrieError(c.KB_E_UNSUPPORTED, "the oracle declines this rule set")
        f = w["facts"]
        if len(f):  # infer_generic.rs:46: inferred facts join the store
            self.store_append(f[:, 0], f[:, 1], f[:, 2], 0)
        return self.rel_from_host([0, 1, 2], [f[:, k].copy() for k in range(3)]), OracleCtx._Stats(w)

    def datalog_fixpoint_seed(self, rules, seed, strategy=0):
        """accepted = seed facts (of rule predicates) the store does not hold; inferred = closure(store + accepted) beyond that"""
        sd = seed.to_numpy([0, 1, 2])
        preds = {int(x.p.value) for r in rules for x in list(r["premise"]) + list(r["conclusion"]) if not x.p.is_var}
        sd = sd[np.isin(sd[:, 1], np.fromiter(preds, np.uint32, len(preds)))] if len(sd) else sd
        have = set(map(tuple, self._rows().tolist()))
        acc = np.array([t for t in dict.fromkeys(map(tuple, sd.tolist())) if t not in have], np.uint32).reshape(-1, 3)
        if len(acc):
            self.store_append(acc[:, 0], acc[:, 1], acc[:, 2], 0)
        rel, st = self.datalog_fixpoint(rules, strategy) if len(acc) else (self.rel_from_host([0, 1, 2], [np.empty(0, np.uint32)] * 3), OracleCtx._Stats())
        out = np.concatenate([acc, rel.to_numpy([0, 1, 2])], axis=0)
        return self.rel_from_host([0, 1, 2], [out[:, k].copy() for k in range(3)]), len(acc), st

    def close(self):
        self.db = None
