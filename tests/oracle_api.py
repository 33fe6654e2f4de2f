"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (the checker), never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from kolibrie_b200 import capi as c  # struct layouts of the boundary header only

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
_L = None


def usable_cpus() -> int:
    """CPUs this process may really use: affinity mask and cgroup quota, not the machine's thread count. An OpenMP team sized by
    `nproc` on a box whose container is capped far below it spends its time in oversubscribed barriers (measured: 0.4-1.1 s per tiny
    oracle query with 128 threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(q) // int(per)))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except (OSError, ValueError):
            pass
    return max(1, n)


def lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB):
            raise ImportError(f"{LIB} missing: run `make -C oracle`")
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # idle team members sleep instead of spinning next to the CUDA host thread
        L = C.CDLL(LIB)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        P = C.POINTER
        sig = {
            "ko_db_create": (vp, [vp, vp, vp, u64]), "ko_db_numeric": (None, [vp, vp, vp, u32]), "ko_db_build_index": (None, [vp]),
            "ko_db_free": (None, [vp]),
            "ko_bgp_execute": (vp, [vp, C.c_int, P(c.KbPattern), u32, P(c.KbFilterOp), u32, P(u32), u32]),
            "ko_scan": (vp, [vp, P(c.KbPattern), P(c.KbFilterOp), u32]),
            "ko_rel_from_host": (vp, [P(u32), u32, P(vp), u64]),
            "ko_hash_join": (vp, [vp, vp]), "ko_filter": (vp, [vp, vp, P(c.KbFilterOp), u32]), "ko_project": (vp, [vp, P(u32), u32]),
            "ko_rel_rows": (u64, [vp]), "ko_rel_cols": (u32, [vp]), "ko_rel_slot": (u32, [vp, u32]), "ko_rel_col": (P(u32), [vp, u32]),
            "ko_rel_free": (None, [vp]),
            "ko_group_aggregate": (vp, [vp, vp, P(u32), u32, P(c.KbAgg), u32]),
            "ko_groups_n": (u64, [vp]), "ko_groups_keys": (P(u32), [vp, u32]), "ko_groups_values": (P(C.c_double), [vp, u32]),
            "ko_groups_counts": (P(u64), [vp]), "ko_groups_free": (None, [vp]),
            "ko_datalog_fixpoint": (vp, [vp, P(c.KbRule), u32, u32]),
            "ko_fix_status": (C.c_int, [vp]), "ko_fix_n": (u64, [vp]), "ko_fix_copy": (None, [vp, vp, vp, vp]), "ko_fix_rounds": (u32, [vp]),
            "ko_fix_round_new": (u64, [vp, u32]), "ko_fix_derivations": (u64, [vp]), "ko_fix_free": (None, [vp]),
            "ko_legacy_select": (u32, [vp, vp, u32, u32, P(u32), vp]),
            "ko_rust_parse_f64": (C.c_int, [C.c_char_p, u64, P(C.c_double)]),
            "ko_num_threads": (C.c_int, []), "ko_set_threads": (None, [C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _L = L
        L.ko_set_threads(min(usable_cpus(), int(os.environ.get("KOLIBRIE_ORACLE_THREADS", "16"))))  # checker runs: small inputs
    return _L


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Rel:
    def __init__(self, h):
        self.h = C.c_void_p(h)

    @property
    def n_rows(self):
        return lib().ko_rel_rows(self.h)

    @property
    def slots(self):
        return [lib().ko_rel_slot(self.h, i) for i in range(lib().ko_rel_cols(self.h))]

    def to_numpy(self, slot_order: Optional[Sequence[int]] = None):
        n, slots = self.n_rows, self.slots
        order = list(slot_order) if slot_order is not None else slots
        out = np.empty((n, len(order)), dtype=np.uint32)
        for j, s in enumerate(order):
            if n:
                out[:, j] = np.ctypeslib.as_array(lib().ko_rel_col(self.h, slots.index(s)), shape=(n,))
        return out

    def __del__(self):
        try:
            if self.h:
                lib().ko_rel_free(self.h)
                self.h = None
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass


class Db:
    def __init__(self, s, p, o, num_or0=None, is_num=None):
        s, p, o = _u32(s), _u32(p), _u32(o)
        self.h = C.c_void_p(lib().ko_db_create(s.ctypes.data, p.ctypes.data, o.ctypes.data, len(s)))
        if num_or0 is not None:
            num = np.ascontiguousarray(num_or0, dtype=np.float64)
            isn = np.ascontiguousarray(is_num, dtype=np.uint8)
            lib().ko_db_numeric(self.h, num.ctypes.data, isn.ctypes.data, len(num))

    def build_index(self):
        lib().ko_db_build_index(self.h)

    def bgp(self, pats, filt=None, project=None, mode=0) -> Rel:
        a, n = c.filter_prog(filt)
        if project is not None:
            pr = (C.c_uint32 * max(len(project), 1))(*project)
            return Rel(lib().ko_bgp_execute(self.h, mode, c.patterns(pats), len(pats), a, n, pr, len(project)))
        return Rel(lib().ko_bgp_execute(self.h, mode, c.patterns(pats), len(pats), a, n, None, 0))

    def scan(self, pat, filt=None) -> Rel:
        a, n = c.filter_prog(filt)
        return Rel(lib().ko_scan(self.h, C.byref(pat), a, n))

    def filter(self, rel: Rel, filt) -> Rel:
        a, n = c.filter_prog(filt)
        return Rel(lib().ko_filter(self.h, rel.h, a, n))

    def group(self, rel: Rel, group_slots, aggs):
        gs = (C.c_uint32 * max(len(group_slots), 1))(*group_slots)
        ag = (c.KbAgg * max(len(aggs), 1))()
        for i, (k, s) in enumerate(aggs):
            ag[i] = c.KbAgg(k, s)
        g = C.c_void_p(lib().ko_group_aggregate(self.h, rel.h, gs, len(group_slots), ag, len(aggs)))
        n = lib().ko_groups_n(g)
        keys = [np.ctypeslib.as_array(lib().ko_groups_keys(g, i), shape=(n,)).copy() if n else np.empty(0, np.uint32) for i in range(len(group_slots))]
        vals = [np.ctypeslib.as_array(lib().ko_groups_values(g, i), shape=(n,)).copy() if n else np.empty(0) for i in range(len(aggs))]
        counts = np.ctypeslib.as_array(lib().ko_groups_counts(g), shape=(n,)).copy() if n else np.empty(0, np.uint64)
        lib().ko_groups_free(g)
        return {"keys": keys, "values": vals, "counts": counts}

    def fixpoint(self, rules, strategy=0):
        arr, keep = c.make_rules(rules)
        f = C.c_void_p(lib().ko_datalog_fixpoint(self.h, arr, len(rules), strategy))
        try:
            st = lib().ko_fix_status(f)
            n = lib().ko_fix_n(f)
            s, p, o = (np.empty(n, dtype=np.uint32) for _ in range(3))
            lib().ko_fix_copy(f, s.ctypes.data, p.ctypes.data, o.ctypes.data)
            rounds = [lib().ko_fix_round_new(f, r) for r in range(lib().ko_fix_rounds(f))]
            return {"status": st, "facts": np.stack([s, p, o], axis=1) if n else np.empty((0, 3), np.uint32), "round_new": rounds,
                    "derivations": lib().ko_fix_derivations(f)}
        finally:
            lib().ko_fix_free(f)

    def __del__(self):
        try:
            if self.h:
                lib().ko_db_free(self.h)
                self.h = None
        except Exception:
            pass


def rel_from_host(slots, cols) -> Rel:
    cols = [_u32(x) for x in cols]
    n = len(cols[0]) if cols else 0
    sl = (C.c_uint32 * max(len(slots), 1))(*slots)
    ptrs = (C.c_void_p * max(len(cols), 1))(*[x.ctypes.data for x in cols])
    r = Rel(lib().ko_rel_from_host(sl, len(slots), ptrs, n))
    r._keep = cols
    return r


def hash_join(a: Rel, b: Rel) -> Rel:
    return Rel(lib().ko_hash_join(a.h, b.h))


def project(a: Rel, slots) -> Rel:
    sl = (C.c_uint32 * max(len(slots), 1))(*slots)
    return Rel(lib().ko_project(a.h, sl, len(slots)))


def legacy_select(p, o, pred, literal=None):
    p, o = _u32(p), _u32(o)
    out = np.empty(len(p), dtype=np.uint32)
    lit = C.c_uint32(literal) if literal is not None else None
    n = lib().ko_legacy_select(p.ctypes.data, o.ctypes.data, len(p), pred, C.byref(lit) if lit is not None else None, out.ctypes.data)
    return out[:n].copy()


def rust_parse_f64(s: str):
    b = s.encode()
    out = C.c_double()
    ok = lib().ko_rust_parse_f64(b, len(b), C.byref(out))
    return out.value if ok else None


def num_threads():
    return lib().ko_num_threads()


def set_threads(n):
    lib().ko_set_threads(n)
