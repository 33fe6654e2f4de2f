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

    def close(self):
        self.db = None
