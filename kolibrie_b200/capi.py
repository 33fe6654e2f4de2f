"""ctypes binding of libkolibrie_b200.so (include/kolibrie_b200.h).

This is the Python stand-in for the Rust shim described in INTEGRATION.md: the reference is a Rust workspace and there is no
Rust toolchain in this image, so tests, bench.py and the multi-GPU plumbing drive the C ABI from Python. There is no CPU
fallback: if the library (or a GPU) is missing, every entry point raises.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref
from typing import List, Iterable, Optional, Sequence, Tuple, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KOLIBRIE_B200_LIB") or os.path.join(_HERE, "libkolibrie_b200.so")  # override: A/B of build variants
LEGACY_LIB_PATH = os.path.join(_HERE, "libcudajoin.so")

KB_OK = 0
KB_E_INVALID, KB_E_CUDA, KB_E_OOM, KB_E_UNSUPPORTED, KB_E_NOT_FOUND, KB_E_LIMIT = -1, -2, -3, -4, -5, -6
KB_ID_NONE = 0xFFFFFFFF
KB_MAX_COLS = 16
KB_TAG_INFERRED = 0xFFFFFFFFFFFFFFF0

# filter opcodes / comparisons / aggregates (mirror the header)
F_CMP_NUM, F_EQ_ID, F_NE_ID, F_AND, F_OR, F_NOT, F_PUSH_VAR, F_PUSH_CONST, F_ADD, F_SUB, F_MUL, F_DIV, F_TRUTHY, F_IS_TRIPLE, F_CMP_LEGACY = range(1, 16)
LEGACY_CONST_IS_I32 = 0x100
LEGACY_NESTED_F64 = 0x200
CMP_GT, CMP_GE, CMP_LT, CMP_LE, CMP_EQ, CMP_NE = range(1, 7)
AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX, AGG_AVG = range(5)
SEMI_NAIVE, NAIVE, SEMI_NAIVE_PARALLEL = 0, 1, 2
SEMI_NAIVE_OLD_DELTA = 3  # textbook OLD/delta scheme: same facts, rounds and per-round counts, fewer candidates (include/kolibrie_b200.h)


class KbTerm(C.Structure):
    _fields_ = [("is_var", C.c_uint32), ("value", C.c_uint32)]


class KbPattern(C.Structure):
    _fields_ = [("s", KbTerm), ("p", KbTerm), ("o", KbTerm)]


class KbFilterOp(C.Structure):
    _fields_ = [("op", C.c_uint32), ("slot", C.c_uint32), ("cmp", C.c_uint32), ("id", C.c_uint32), ("value", C.c_double)]


class KbAgg(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("slot", C.c_uint32)]


class KbRuleFilter(C.Structure):
    _fields_ = [("lhs_slot", C.c_uint32), ("cmp", C.c_uint32), ("rhs_is_var", C.c_uint32), ("rhs_slot", C.c_uint32), ("rhs_value", C.c_double)]


class KbRule(C.Structure):
    _fields_ = [
        ("premise", C.POINTER(KbPattern)), ("n_premise", C.c_uint32),
        ("filters", C.POINTER(KbRuleFilter)), ("n_filters", C.c_uint32),
        ("conclusion", C.POINTER(KbPattern)), ("n_conclusion", C.c_uint32),
    ]


class KbFixpointStats(C.Structure):
    _fields_ = [("rounds", C.c_uint32), ("inferred", C.c_uint64), ("derivations", C.c_uint64), ("round_new", C.c_uint64 * 64), ("device_ms", C.c_double)]


class KbStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("scan_ms", "build_ms", "probe_ms", "filter_ms", "group_ms", "other_ms", "total_ms")] + [
        (n, C.c_uint64)
        for n in ("scan_launches", "build_launches", "probe_launches", "filter_launches", "group_launches", "other_launches",
                  "rows_scanned", "rows_built", "rows_probed", "rows_out", "h2d_bytes", "d2h_bytes", "kernel_launches", "fused_scan_builds", "index_joins")
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def V(slot: int) -> KbTerm:
    """variable term"""
    return KbTerm(1, slot)


def K(term_id: int) -> KbTerm:
    """constant term"""
    return KbTerm(0, term_id)


def pattern(s: KbTerm, p: KbTerm, o: KbTerm) -> KbPattern:
    return KbPattern(s, p, o)


def patterns(pats: Sequence[KbPattern]):
    arr = (KbPattern * max(len(pats), 1))()
    for i, p in enumerate(pats):
        arr[i] = p
    return arr


def fop(op, slot=0, cmp=0, id=0, value=0.0) -> KbFilterOp:
    return KbFilterOp(op, slot, cmp, id, value)


def filter_prog(ops: Optional[Sequence[KbFilterOp]]):
    ops = list(ops or [])
    arr = (KbFilterOp * max(len(ops), 1))()
    for i, o in enumerate(ops):
        arr[i] = o
    return arr, len(ops)


class KolibrieError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"kb_status {status}: {message}")
        self.status = status
        self.message = message


_lib = None
_LIVE_CONTEXTS = weakref.WeakSet()


@atexit.register
def _close_all():
    for ctx in list(_LIVE_CONTEXTS):
        try:
            ctx.close()
        except Exception:
            pass


def lib() -> C.CDLL:
    """Load the shared library; fail loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                          "kolibrie_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    sig = {
        "kb_version": (C.c_char_p, []),
        "kb_ctx_create": (i32, [C.c_int, P(vp)]),
        "kb_ctx_destroy": (None, [vp]),
        "kb_last_error": (C.c_char_p, [vp]),
        "kb_set_timing": (i32, [vp, C.c_int]),
        "kb_get_stats": (i32, [vp, P(KbStats), C.c_int]),
        "kb_synchronize": (i32, [vp]),
        "kb_store_load": (i32, [vp, vp, vp, vp, u64]),
        "kb_store_load_device": (i32, [vp, vp, vp, vp, u64]),
        "kb_store_append_device": (i32, [vp, vp, vp, vp, u64, u64]),
        "kb_store_append": (i32, [vp, vp, vp, vp, u64, u64]),
        "kb_store_evict": (i32, [vp, u64]),
        "kb_store_delete": (i32, [vp, vp, vp, vp, u64]),
        "kb_store_clear": (i32, [vp]),
        "kb_store_build_index": (i32, [vp, P(u32), P(C.c_double)]),
        "kb_set_use_index": (i32, [vp, C.c_int]),
        "kb_store_size": (i32, [vp, P(u64), P(u32)]),
        "kb_store_download": (i32, [vp, vp, vp, vp, u64, P(u64)]),
        "kb_dict_numeric_load": (i32, [vp, vp, vp, u32]),
        "kb_dict_strings_load": (i32, [vp, vp, vp, u32]),
        "kb_dict_encode": (i32, [vp, vp, vp, u64, vp, P(u32), vp]),
        "kb_dict_strings_info": (i32, [vp, P(u32), P(u64)]),
        "kb_dict_legacy_i32_load": (i32, [vp, vp, vp, u32]),
        "kb_rel_decode": (i32, [vp, vp, u32, P(vp)]),
        "kb_strings_info": (i32, [vp, P(u64), P(u64)]),
        "kb_strings_download": (i32, [vp, vp, vp, vp]),
        "kb_strings_free": (None, [vp, vp]),
        "kb_rel_info": (i32, [vp, P(u64), P(u32), P(u32)]),
        "kb_rel_download": (i32, [vp, vp, u32, vp]),
        "kb_rel_device_col": (i32, [vp, u32, P(vp)]),
        "kb_rel_from_host": (i32, [vp, P(u32), u32, P(vp), u64, P(vp)]),
        "kb_rel_from_device": (i32, [vp, P(u32), u32, P(vp), u64, P(vp)]),
        "kb_rel_free": (None, [vp, vp]),
        "kb_scan": (i32, [vp, P(KbPattern), u32, P(P(KbFilterOp)), P(u32), P(vp)]),
        "kb_filter": (i32, [vp, vp, P(KbFilterOp), u32, P(vp)]),
        "kb_project": (i32, [vp, vp, P(u32), u32, P(vp)]),
        "kb_hash_join": (i32, [vp, vp, vp, P(vp)]),
        "kb_bind_join": (i32, [vp, vp, P(KbPattern), P(vp)]),
        "kb_star_join": (i32, [vp, u32, P(KbPattern), u32, P(KbFilterOp), u32, P(vp)]),
        "kb_bgp_execute": (i32, [vp, P(KbPattern), u32, P(KbFilterOp), u32, P(u32), u32, P(vp)]),
        "kb_group_aggregate": (i32, [vp, vp, P(u32), u32, P(KbAgg), u32, P(vp)]),
        "kb_groups_info": (i32, [vp, P(u64), P(u32), P(u32)]),
        "kb_groups_keys": (i32, [vp, u32, P(P(u32))]),
        "kb_groups_values": (i32, [vp, u32, P(P(C.c_double))]),
        "kb_groups_counts": (i32, [vp, P(P(u64))]),
        "kb_groups_free": (None, [vp]),
        "kb_groups_pack": (i32, [vp, vp, u64, P(u64)]),
        "kb_groups_merge": (i32, [vp, P(vp), P(u64), u32, P(vp)]),
        "kb_star_join_prepare": (i32, [vp, u32, P(KbPattern), u32, P(KbFilterOp), u32, P(u32), u32, P(KbAgg), u32, u32, P(vp)]),
        "kb_plan_submit": (i32, [vp, vp, P(u64)]),
        "kb_plan_collect": (i32, [vp, vp, u64, P(u64), P(vp), P(vp)]),
        "kb_plan_info": (i32, [vp, P(u32), P(u64), P(u32), P(u32), P(u32)]),
        "kb_plan_free": (None, [vp, vp]),
        "kb_plan_peer_scratch_bytes": (u64, [vp]),
        "kb_plan_attach_peers": (i32, [vp, vp, u32, u32, P(vp)]),
        "kb_star_join_aggregate": (i32, [vp, u32, P(KbPattern), u32, P(KbFilterOp), u32, P(u32), u32, P(KbAgg), u32, P(vp), P(u64)]),
        "kb_datalog_fixpoint": (i32, [vp, P(KbRule), u32, u32, P(vp), P(KbFixpointStats)]),
        "kb_datalog_fixpoint_seed": (i32, [vp, P(KbRule), u32, u32, vp, P(vp), P(u64), P(KbFixpointStats)]),
        "kb_shard_of": (u32, [u32, u32]),
        "kb_set_sharding": (i32, [vp, u32, u32]),
        "kb_partition": (i32, [vp, vp, u32, u32, P(vp), P(u64)]),
        "kb_partition_counts": (i32, [vp, vp, u32, u32, P(u64)]),
        "kb_shuffle_scatter": (i32, [vp, vp, u32, u32, P(vp), P(u64), u64]),
        "kb_segment_write": (i32, [C.c_char_p, vp, vp, vp, u64, u64]),
        "kb_segment_info": (i32, [C.c_char_p, P(u64), P(u64), P(u32), P(u32)]),
        "kb_segment_save": (i32, [vp, u64, C.c_int, C.c_char_p]),
        "kb_store_append_file": (i32, [vp, C.c_char_p, u64, C.c_int]),
        "kb_shuffle_push": (i32, [vp, vp, u32, u32, P(vp), P(vp), u64]),
        "kb_rel_wrap_device": (i32, [vp, P(u32), u32, P(vp), u64, P(vp)]),
        "kb_star_join_host": (i32, [vp, vp, vp, vp, u64, u32, P(KbPattern), u32, P(KbFilterOp), u32, P(u32), P(u32), P(vp), P(u64)]),
        "kb_star_join_host_into": (i32, [vp, vp, vp, vp, u64, u32, P(KbPattern), u32, P(KbFilterOp), u32, P(u32), P(u32), P(vp), u64, P(u64)]),
        "perform_hash_join_cuda": (None, [vp, vp, vp, u32, u32, P(u32), P(P(u32)), P(u32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "kb_version", "kb_ctx_create", "kb_ctx_destroy", "kb_last_error", "kb_set_timing", "kb_get_stats", "kb_synchronize",
    "kb_store_load", "kb_store_load_device", "kb_store_append", "kb_store_append_device", "kb_store_evict", "kb_store_delete", "kb_store_clear", "kb_store_build_index", "kb_set_use_index", "kb_store_size",
    "kb_store_download", "kb_dict_numeric_load", "kb_dict_legacy_i32_load", "kb_dict_strings_load", "kb_dict_encode", "kb_dict_strings_info", "kb_rel_decode", "kb_strings_info", "kb_strings_download", "kb_strings_free", "kb_rel_info", "kb_rel_download", "kb_rel_device_col", "kb_rel_from_host",
    "kb_rel_from_device", "kb_rel_free", "kb_scan", "kb_filter", "kb_project", "kb_hash_join", "kb_bind_join", "kb_star_join", "kb_bgp_execute",
    "kb_group_aggregate", "kb_groups_info", "kb_groups_keys", "kb_groups_values", "kb_groups_counts", "kb_groups_free", "kb_groups_pack", "kb_groups_merge", "kb_star_join_aggregate",
    "kb_star_join_prepare", "kb_plan_submit", "kb_plan_collect", "kb_plan_info", "kb_plan_free", "kb_plan_peer_scratch_bytes", "kb_plan_attach_peers",
    "kb_datalog_fixpoint", "kb_datalog_fixpoint_seed", "kb_shard_of", "kb_set_sharding", "kb_partition", "kb_partition_counts", "kb_segment_write", "kb_segment_info", "kb_segment_save", "kb_store_append_file", "kb_shuffle_scatter", "kb_shuffle_push", "kb_rel_wrap_device", "kb_star_join_host", "kb_star_join_host_into", "perform_hash_join_cuda",
]


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Relation:
    """Device-resident columnar bag of binding rows (kb_rel)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx = ctx
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        ctx._rels.add(self)

    def info(self):
        n, nc = C.c_uint64(), C.c_uint32()
        slots = (C.c_uint32 * KB_MAX_COLS)()
        self.ctx._check(lib().kb_rel_info(self.h, C.byref(n), C.byref(nc), slots))
        return n.value, [slots[i] for i in range(nc.value)]

    @property
    def n_rows(self) -> int:
        return self.info()[0]

    @property
    def slots(self):
        return self.info()[1]

    def column(self, col: int) -> np.ndarray:
        n, _ = self.info()
        out = np.empty(n, dtype=np.uint32)
        self.ctx._check(lib().kb_rel_download(self.ctx.h, self.h, col, _ptr(out)))
        return out

    def decode(self, col: int):
        """kb_rel_decode: the column's ids as strings, decoded on the device. Returns (offsets uint64 [n+1], bytes uint8): string i
        is bytes[offsets[i]:offsets[i+1]] (UTF-8); ids the dictionary does not hold read "unknown" (engine.rs:44)."""
        h = C.c_void_p()
        self.ctx._check(lib().kb_rel_decode(self.ctx.h, self.h, col, C.byref(h)))
        try:
            n, total = C.c_uint64(), C.c_uint64()
            self.ctx._check(lib().kb_strings_info(h, C.byref(n), C.byref(total)))
            off = np.empty(n.value + 1, dtype=np.uint64)
            data = np.empty(total.value, dtype=np.uint8)
            self.ctx._check(lib().kb_strings_download(self.ctx.h, h, _ptr(off), _ptr(data)))
        finally:
            lib().kb_strings_free(self.ctx.h, h)
        return off, data

    def decode_strings(self, col: int) -> List[str]:
        off, data = self.decode(col)
        raw = data.tobytes()
        return [raw[int(off[i]):int(off[i + 1])].decode("utf-8") for i in range(len(off) - 1)]

    def device_ptr(self, col: int) -> int:
        p = C.c_void_p()
        self.ctx._check(lib().kb_rel_device_col(self.h, col, C.byref(p)))
        return p.value or 0

    def to_numpy(self, slot_order: Optional[Sequence[int]] = None) -> np.ndarray:
        """rows x cols uint32 array, columns ordered by `slot_order` (default: the relation's own order)"""
        n, slots = self.info()
        order = list(slot_order) if slot_order is not None else slots
        out = np.empty((n, len(order)), dtype=np.uint32)
        for j, s in enumerate(order):
            out[:, j] = self.column(slots.index(s))
        return out

    def free(self):
        if self.h:
            lib().kb_rel_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One device context = one GPU (multi-GPU = one process per GPU, see kolibrie_b200.dist)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib().kb_ctx_create(device, C.byref(h))
        if rc != KB_OK:
            raise KolibrieError(rc, (lib().kb_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self._rels = weakref.WeakSet()
        self._plans = weakref.WeakSet()
        _LIVE_CONTEXTS.add(self)

    def close(self):
        if self.h:
            for pl in list(self._plans):
                pl.free()
            for r in list(self._rels):  # relations first, while their stream still exists
                r.free()
            lib().kb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != KB_OK:
            raise KolibrieError(rc, (lib().kb_last_error(self.h) or b"").decode())

    # ---- store
    def store_load(self, s, p, o):
        s, p, o = _u32(s), _u32(p), _u32(o)
        self._check(lib().kb_store_load(self.h, _ptr(s), _ptr(p), _ptr(o), len(s)))

    def store_load_device(self, d_s: int, d_p: int, d_o: int, n: int):
        self._check(lib().kb_store_load_device(self.h, C.c_void_p(d_s), C.c_void_p(d_p), C.c_void_p(d_o), n))

    def store_append(self, s, p, o, tag: int):
        s, p, o = _u32(s), _u32(p), _u32(o)
        self._check(lib().kb_store_append(self.h, _ptr(s), _ptr(p), _ptr(o), len(s), tag))

    def store_append_device(self, d_s: int, d_p: int, d_o: int, n: int, tag: int):
        """a slide whose columns are already in this device's memory (copied device-to-device; index maintained as store_append)"""
        self._check(lib().kb_store_append_device(self.h, C.c_void_p(d_s), C.c_void_p(d_p), C.c_void_p(d_o), n, tag))

    def store_evict(self, tag: int):
        self._check(lib().kb_store_evict(self.h, tag))

    def store_delete(self, s, p, o):
        s, p, o = _u32(s), _u32(p), _u32(o)
        self._check(lib().kb_store_delete(self.h, _ptr(s), _ptr(p), _ptr(o), len(s)))

    def store_clear(self):
        self._check(lib().kb_store_clear(self.h))

    def segment_save(self, path: str, tag: Optional[int] = None):
        """kb_segment_save: the segment(s) tagged `tag` — or the whole store when tag is None — as one columnar segment file"""
        self._check(lib().kb_segment_save(self.h, 0 if tag is None else tag, 1 if tag is None else 0, path.encode()))

    def store_append_file(self, path: str, tag: int, verify: bool = True):
        """kb_store_append_file: stream a columnar segment file into a new store segment"""
        self._check(lib().kb_store_append_file(self.h, path.encode(), tag, 1 if verify else 0))

    def build_index(self):
        """SparqlDatabase::build_all_indexes on the device: partition the store by predicate. Returns (n_predicates, build_ms)."""
        n, ms = C.c_uint32(), C.c_double()
        self._check(lib().kb_store_build_index(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def set_use_index(self, on: bool):
        self._check(lib().kb_set_use_index(self.h, 1 if on else 0))

    def store_size(self):
        n, ns = C.c_uint64(), C.c_uint32()
        self._check(lib().kb_store_size(self.h, C.byref(n), C.byref(ns)))
        return n.value, ns.value

    def store_download(self):
        n, _ = self.store_size()
        s, p, o = (np.empty(n, dtype=np.uint32) for _ in range(3))
        got = C.c_uint64()
        self._check(lib().kb_store_download(self.h, _ptr(s), _ptr(p), _ptr(o), n, C.byref(got)))
        return s, p, o

    def dict_numeric_load(self, num_or0, is_num):
        num = np.ascontiguousarray(num_or0, dtype=np.float64)
        isn = np.ascontiguousarray(is_num, dtype=np.uint8)
        assert len(num) == len(isn)
        self._check(lib().kb_dict_numeric_load(self.h, _ptr(num), _ptr(isn), len(num)))

    def dict_legacy_i32_load(self, val, is_i32):
        """kb_dict_legacy_i32_load: the i32 view of the terms the legacy executor's FILTER compares with (apply_filters_simd)"""
        v = np.ascontiguousarray(val, dtype=np.int32)
        f = np.ascontiguousarray(is_i32, dtype=np.uint8)
        assert len(v) == len(f)
        self._check(lib().kb_dict_legacy_i32_load(self.h, _ptr(v), _ptr(f), len(v)))

    def dict_strings_load(self, strings: Sequence[str]):
        """kb_dict_strings_load: string i = the term with dictionary id i (Dictionary::id_to_string, dictionary.rs:17-21)"""
        enc = [x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in strings]
        off = np.zeros(len(enc) + 1, dtype=np.uint64)
        if enc:
            off[1:] = np.cumsum([len(b) for b in enc], dtype=np.uint64)
        data = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc else np.empty(0, np.uint8)
        data = np.ascontiguousarray(data)
        self._check(lib().kb_dict_strings_load(self.h, _ptr(off), _ptr(data) if len(data) else None, len(enc)))

    def dict_encode(self, terms: Sequence[Union[str, bytes]]):
        """kb_dict_encode: Dictionary::encode of a batch of terms on the device. Returns (ids uint32[n], first_pos uint64[n_new]):
        first_pos[k] = index in `terms` of the term that introduced id (ids before the call) + k."""
        enc = [x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in terms]
        n = len(enc)
        ids = np.empty(n, dtype=np.uint32)
        if n == 0:
            return ids, np.empty(0, np.uint64)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.uint64, count=n), dtype=np.uint64)
        data = np.ascontiguousarray(np.frombuffer(b"".join(enc), dtype=np.uint8)) if off[-1] else np.empty(0, np.uint8)
        first = np.empty(n, dtype=np.uint64)
        n_new = C.c_uint32(0)
        self._check(lib().kb_dict_encode(self.h, _ptr(off), _ptr(data) if len(data) else None, n, _ptr(ids), C.byref(n_new), _ptr(first)))
        return ids, first[: n_new.value].copy()

    def dict_strings_info(self) -> Tuple[int, int]:
        n, b = C.c_uint32(0), C.c_uint64(0)
        self._check(lib().kb_dict_strings_info(self.h, C.byref(n), C.byref(b)))
        return int(n.value), int(b.value)

    # ---- operators
    def scan(self, pats: Sequence[KbPattern], pushdown: Optional[Sequence[Optional[Sequence[KbFilterOp]]]] = None):
        arr = patterns(pats)
        k = len(pats)
        out = (C.c_void_p * k)()
        if pushdown:
            progs, keep = (C.POINTER(KbFilterOp) * k)(), []
            lens = (C.c_uint32 * k)()
            for i in range(k):
                a, n = filter_prog(pushdown[i] if i < len(pushdown) else None)
                keep.append(a)
                progs[i] = C.cast(a, C.POINTER(KbFilterOp))
                lens[i] = n
            self._check(lib().kb_scan(self.h, arr, k, progs, lens, out))
        else:
            self._check(lib().kb_scan(self.h, arr, k, None, None, out))
        return [Relation(self, out[i]) for i in range(k)]

    def filter(self, rel: Relation, ops: Sequence[KbFilterOp]) -> Relation:
        a, n = filter_prog(ops)
        out = C.c_void_p()
        self._check(lib().kb_filter(self.h, rel.h, a, n, C.byref(out)))
        return Relation(self, out)

    def project(self, rel: Relation, slots: Sequence[int]) -> Relation:
        a = (C.c_uint32 * max(len(slots), 1))(*slots)
        out = C.c_void_p()
        self._check(lib().kb_project(self.h, rel.h, a, len(slots), C.byref(out)))
        return Relation(self, out)

    def hash_join(self, left: Relation, right: Relation) -> Relation:
        out = C.c_void_p()
        self._check(lib().kb_hash_join(self.h, left.h, right.h, C.byref(out)))
        return Relation(self, out)

    def bind_join(self, left: Relation, pat: KbPattern) -> Relation:
        """kb_bind_join: `left` joined with one store pattern (index lookup kernel when the index has a table for it)"""
        out = C.c_void_p()
        arr = patterns([pat])
        self._check(lib().kb_bind_join(self.h, left.h, arr, C.byref(out)))
        return Relation(self, out)

    def star_join(self, join_slot: int, pats: Sequence[KbPattern], filt: Optional[Sequence[KbFilterOp]] = None) -> Relation:
        a, n = filter_prog(filt)
        out = C.c_void_p()
        self._check(lib().kb_star_join(self.h, join_slot, patterns(pats), len(pats), a, n, C.byref(out)))
        return Relation(self, out)

    def bgp_execute(self, pats: Sequence[KbPattern], filt: Optional[Sequence[KbFilterOp]] = None, project: Optional[Sequence[int]] = None) -> Relation:
        a, n = filter_prog(filt)
        out = C.c_void_p()
        if project is not None:
            pr = (C.c_uint32 * max(len(project), 1))(*project)
            self._check(lib().kb_bgp_execute(self.h, patterns(pats), len(pats), a, n, pr, len(project), C.byref(out)))
        else:
            self._check(lib().kb_bgp_execute(self.h, patterns(pats), len(pats), a, n, None, 0, C.byref(out)))
        return Relation(self, out)

    def rel_from_host(self, slots: Sequence[int], cols: Sequence[np.ndarray]) -> Relation:
        cols = [_u32(c) for c in cols]
        n = len(cols[0]) if cols else 0
        sl = (C.c_uint32 * max(len(slots), 1))(*slots)
        ptrs = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        out = C.c_void_p()
        self._check(lib().kb_rel_from_host(self.h, sl, len(slots), ptrs, n, C.byref(out)))
        return Relation(self, out)

    def rel_from_device(self, slots: Sequence[int], ptrs_: Sequence[int], n: int) -> Relation:
        sl = (C.c_uint32 * max(len(slots), 1))(*slots)
        ptrs = (C.c_void_p * max(len(ptrs_), 1))(*ptrs_)
        out = C.c_void_p()
        self._check(lib().kb_rel_from_device(self.h, sl, len(slots), ptrs, n, C.byref(out)))
        return Relation(self, out)

    def group_aggregate(self, rel: Relation, group_slots: Sequence[int], aggs: Sequence[tuple]):
        """aggs: [(kind, slot)] -> dict(keys=[cols], values=[cols], counts=array)"""
        gs = (C.c_uint32 * max(len(group_slots), 1))(*group_slots)
        ag = (KbAgg * max(len(aggs), 1))()
        for i, (k, s) in enumerate(aggs):
            ag[i] = KbAgg(k, s)
        g = C.c_void_p()
        self._check(lib().kb_group_aggregate(self.h, rel.h, gs, len(group_slots), ag, len(aggs), C.byref(g)))
        return self._groups_to_dict(g)

    def star_join_aggregate(self, join_slot: int, pats: Sequence[KbPattern], filt, group_slots: Sequence[int], aggs: Sequence[tuple]):
        """kb_star_join_aggregate: star join + GROUP BY without materialising the join when the fused path applies.
        Returns (groups dict as group_aggregate, joined row count)."""
        a, nf = filter_prog(filt)
        gs = (C.c_uint32 * max(len(group_slots), 1))(*group_slots)
        ag = (KbAgg * max(len(aggs), 1))()
        for i, (k, s) in enumerate(aggs):
            ag[i] = KbAgg(k, s)
        g = C.c_void_p()
        n_rows = C.c_uint64()
        self._check(lib().kb_star_join_aggregate(self.h, join_slot, patterns(pats), len(pats), a, nf, gs, len(group_slots), ag, len(aggs), C.byref(g), C.byref(n_rows)))
        return self._groups_to_dict(g), n_rows.value

    def star_join_aggregate_packed(self, join_slot: int, pats: Sequence[KbPattern], filt, group_slots: Sequence[int], aggs: Sequence[tuple]):
        """kb_star_join_aggregate, result left in its transport form (kb_groups_pack): what one rank contributes to a cross-rank
        GROUP BY. Returns (uint8 array, joined row count)."""
        a, nf = filter_prog(filt)
        gs = (C.c_uint32 * max(len(group_slots), 1))(*group_slots)
        ag = (KbAgg * max(len(aggs), 1))()
        for i, (k, s) in enumerate(aggs):
            ag[i] = KbAgg(k, s)
        g = C.c_void_p()
        n_rows = C.c_uint64()
        self._check(lib().kb_star_join_aggregate(self.h, join_slot, patterns(pats), len(pats), a, nf, gs, len(group_slots), ag, len(aggs), C.byref(g), C.byref(n_rows)))
        try:
            return self._groups_pack(g), n_rows.value
        finally:
            lib().kb_groups_free(g)

    def _groups_pack(self, g) -> np.ndarray:
        need = C.c_uint64()
        self._check(lib().kb_groups_pack(g, None, 0, C.byref(need)))
        buf = np.empty(need.value, dtype=np.uint8)
        self._check(lib().kb_groups_pack(g, _ptr(buf), need.value, C.byref(need)))
        return buf

    def groups_merge(self, parts: Sequence[np.ndarray]):
        """kb_groups_merge: fold the packed partial GROUP BY results of several ranks (device kernel); returns the groups dict"""
        parts = [np.ascontiguousarray(p_, dtype=np.uint8) for p_ in parts]
        ptrs = (C.c_void_p * len(parts))(*[p_.ctypes.data for p_ in parts])
        sizes = (C.c_uint64 * len(parts))(*[p_.nbytes for p_ in parts])
        g = C.c_void_p()
        self._check(lib().kb_groups_merge(self.h, ptrs, sizes, len(parts), C.byref(g)))
        return self._groups_to_dict(g)

    def prepare_star_join(self, join_slot: int, pats: Sequence[KbPattern], filt=None, group_slots: Sequence[int] = (), aggs: Sequence[tuple] = (), ring: int = 4) -> "Plan":
        """kb_star_join_prepare: resolve the query once; Plan.submit() is then one asynchronous kernel launch"""
        a, nf = filter_prog(filt)
        gs = (C.c_uint32 * max(len(group_slots), 1))(*group_slots)
        ag = (KbAgg * max(len(aggs), 1))()
        for i, (k, s) in enumerate(aggs):
            ag[i] = KbAgg(k, s)
        h = C.c_void_p()
        self._check(lib().kb_star_join_prepare(self.h, join_slot, patterns(pats), len(pats), a, nf, gs, len(group_slots), ag, len(aggs), ring, C.byref(h)))
        return Plan(self, h)

    def _groups_to_dict(self, g):
        try:
            n, ng, na = C.c_uint64(), C.c_uint32(), C.c_uint32()
            self._check(lib().kb_groups_info(g, C.byref(n), C.byref(ng), C.byref(na)))
            keys, vals = [], []
            for c in range(ng.value):
                p = C.POINTER(C.c_uint32)()
                self._check(lib().kb_groups_keys(g, c, C.byref(p)))
                keys.append(np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint32))
            for a in range(na.value):
                p = C.POINTER(C.c_double)()
                self._check(lib().kb_groups_values(g, a, C.byref(p)))
                vals.append(np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.empty(0, np.float64))
            pc = C.POINTER(C.c_uint64)()
            self._check(lib().kb_groups_counts(g, C.byref(pc)))
            counts = np.ctypeslib.as_array(pc, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint64)
            return {"keys": keys, "values": vals, "counts": counts}
        finally:
            lib().kb_groups_free(g)

    def datalog_fixpoint(self, rules: Sequence[dict], strategy: int = SEMI_NAIVE):
        """rules: [{'premise': [KbPattern], 'conclusion': [KbPattern], 'filters': [KbRuleFilter]}]"""
        arr, keep = make_rules(rules)
        out = C.c_void_p()
        st = KbFixpointStats()
        self._check(lib().kb_datalog_fixpoint(self.h, arr, len(rules), strategy, C.byref(out), C.byref(st)))
        return Relation(self, out), st

    def datalog_fixpoint_seed(self, rules: Sequence[dict], seed: Relation, strategy: int = SEMI_NAIVE):
        """kb_datalog_fixpoint_seed: the store is closed under `rules` already; `seed` (slots 0,1,2 = s,p,o) are facts to add. Returns
        (relation = the accepted seed facts followed by the facts inferred from them, number of accepted seed facts, stats)."""
        arr, keep = make_rules(rules)
        out = C.c_void_p()
        n_new = C.c_uint64()
        st = KbFixpointStats()
        self._check(lib().kb_datalog_fixpoint_seed(self.h, arr, len(rules), strategy, seed.h, C.byref(out), C.byref(n_new), C.byref(st)))
        return Relation(self, out), int(n_new.value), st

    def partition(self, rel: Relation, key_slot: int, n_parts: int):
        out = C.c_void_p()
        offs = (C.c_uint64 * (n_parts + 1))()
        self._check(lib().kb_partition(self.h, rel.h, key_slot, n_parts, C.byref(out), offs))
        return Relation(self, out), [offs[i] for i in range(n_parts + 1)]

    def partition_counts(self, rel: Relation, key_slot: int, n_parts: int) -> List[int]:
        cnt = (C.c_uint64 * n_parts)()
        self._check(lib().kb_partition_counts(self.h, rel.h, key_slot, n_parts, cnt))
        return [int(cnt[i]) for i in range(n_parts)]

    def shuffle_scatter(self, rel: Relation, key_slot: int, n_parts: int, peer_cols: Sequence[int], base: Sequence[int], capacity_rows: int):
        """kb_shuffle_scatter: peer_cols[d * n_cols + c] = address of column c of rank d's receive buffer (peer-mapped)"""
        pc = (C.c_void_p * len(peer_cols))(*peer_cols)
        bs = (C.c_uint64 * n_parts)(*base)
        self._check(lib().kb_shuffle_scatter(self.h, rel.h, key_slot, n_parts, pc, bs, capacity_rows))

    def shuffle_push(self, rel: Relation, key_slot: int, n_parts: int, peer_cols: Sequence[int], peer_cursors: Sequence[int], capacity_rows: int):
        """kb_shuffle_push: like shuffle_scatter, ranges reserved on the receivers' own cursors (peer_cursors[d] = address of rank d's
        u32 cursor in peer-mapped memory): no count pass, no count exchange"""
        pc = (C.c_void_p * len(peer_cols))(*peer_cols)
        cu = (C.c_void_p * n_parts)(*peer_cursors)
        self._check(lib().kb_shuffle_push(self.h, rel.h, key_slot, n_parts, pc, cu, capacity_rows))

    def rel_wrap_device(self, slots: Sequence[int], ptrs_: Sequence[int], n: int) -> Relation:
        """kb_rel_wrap_device: a relation over device columns the caller owns (no copy)"""
        sl = (C.c_uint32 * max(len(slots), 1))(*slots)
        ptrs = (C.c_void_p * max(len(ptrs_), 1))(*ptrs_)
        out = C.c_void_p()
        self._check(lib().kb_rel_wrap_device(self.h, sl, len(slots), ptrs, n, C.byref(out)))
        return Relation(self, out)

    def star_join_host(self, s, p, o, join_slot: int, pats: Sequence[KbPattern], filt=None):
        """One-shot host-buffer call (kb_star_join_host): upload + star join + download; result columns are malloc'd by the
        library and copied into numpy arrays here. Returns (rows x cols uint32 array, slots)."""
        s, p, o = _u32(s), _u32(p), _u32(o)
        a, nf = filter_prog(filt)
        n_cols, n_rows = C.c_uint32(), C.c_uint64(0)
        slots = (C.c_uint32 * KB_MAX_COLS)()
        cols = (C.c_void_p * KB_MAX_COLS)()
        self._check(lib().kb_star_join_host(self.h, _ptr(s), _ptr(p), _ptr(o), len(s), join_slot, patterns(pats), len(pats), a, nf,
                                            C.byref(n_cols), slots, cols, C.byref(n_rows)))
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        out = np.empty((n_rows.value, n_cols.value), dtype=np.uint32)
        for c in range(n_cols.value):
            if n_rows.value:
                out[:, c] = np.ctypeslib.as_array(C.cast(cols[c], C.POINTER(C.c_uint32)), shape=(n_rows.value,))
            libc.free(cols[c])
        return out, [slots[i] for i in range(n_cols.value)]

    # ---- stats
    def set_timing(self, on: bool):
        self._check(lib().kb_set_timing(self.h, 1 if on else 0))

    def get_stats(self, reset: bool = False) -> dict:
        st = KbStats()
        self._check(lib().kb_get_stats(self.h, C.byref(st), 1 if reset else 0))
        return st.as_dict()

    def synchronize(self):
        self._check(lib().kb_synchronize(self.h))

    def set_sharding(self, rank: int, world: int):
        """the store holds shard `rank` of `world` (sharded by subject with kb_shard_of): enables the dense key compaction"""
        self._check(lib().kb_set_sharding(self.h, rank, world))


class Plan:
    """kb_plan: a prepared star join with a ring of result buffers. submit() launches, collect(ticket) waits for that launch."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        ring, cap, nc, grouped = C.c_uint32(), C.c_uint64(), C.c_uint32(), C.c_uint32()
        slots = (C.c_uint32 * KB_MAX_COLS)()
        ctx._check(lib().kb_plan_info(self.h, C.byref(ring), C.byref(cap), C.byref(nc), slots, C.byref(grouped)))
        self.ring, self.capacity_rows, self.grouped = ring.value, cap.value, bool(grouped.value)
        self.slots = [slots[i] for i in range(nc.value)]
        self._ticket = C.c_uint64()
        self._rows = C.c_uint64()
        ctx._plans.add(self)

    def submit(self) -> int:
        rc = lib().kb_plan_submit(self.ctx.h, self.h, C.byref(self._ticket))
        if rc != KB_OK:
            self.ctx._check(rc)
        return self._ticket.value

    def collect(self, ticket: int) -> int:
        """joined row count of the query `ticket` (waits for it)"""
        rc = lib().kb_plan_collect(self.ctx.h, self.h, ticket, C.byref(self._rows), None, None)
        if rc != KB_OK:
            self.ctx._check(rc)
        return self._rows.value

    def collect_rows(self, ticket: int) -> Relation:
        """the result as a Relation VIEW of the ring slot (valid until `ring` further submits)"""
        r = C.c_void_p()
        self.ctx._check(lib().kb_plan_collect(self.ctx.h, self.h, ticket, C.byref(self._rows), C.byref(r), None))
        return Relation(self.ctx, r)

    def collect_groups(self, ticket: int, packed: bool = False):
        """(groups dict — or the kb_groups_pack bytes when packed=True —, joined row count) of a grouped plan"""
        g = C.c_void_p()
        self.ctx._check(lib().kb_plan_collect(self.ctx.h, self.h, ticket, C.byref(self._rows), None, C.byref(g)))
        if packed:
            try:
                return self.ctx._groups_pack(g), self._rows.value
            finally:
                lib().kb_groups_free(g)
        return self.ctx._groups_to_dict(g), self._rows.value

    def peer_scratch_bytes(self) -> int:
        return int(lib().kb_plan_peer_scratch_bytes(self.h))

    def attach_peers(self, rank: int, world: int, peer_scratch: Sequence[int], keep=None):
        """kb_plan_attach_peers: peer_scratch[r] = device-visible address of rank r's zeroed scratch (peer_scratch_bytes() each);
        `keep` = whatever owns that memory (kept alive with the plan)"""
        ptrs = (C.c_void_p * world)(*peer_scratch)
        self.ctx._check(lib().kb_plan_attach_peers(self.ctx.h, self.h, rank, world, ptrs))
        self._keep = keep

    def free(self):
        if self.h:
            lib().kb_plan_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def star_join_host_raw(ctx: Context, s_ptr: int, p_ptr: int, o_ptr: int, n: int, join_slot: int, pats, filt, out_ptrs: Sequence[int], out_cap: int):
    """kb_star_join_host_into with raw host pointers (pinned torch / numpy memory) for inputs and outputs. Returns (n_rows, slots)."""
    a, nf = filter_prog(filt)
    n_cols, n_rows = C.c_uint32(), C.c_uint64(0)
    slots = (C.c_uint32 * KB_MAX_COLS)()
    cols = (C.c_void_p * KB_MAX_COLS)()
    for i, ptr in enumerate(out_ptrs):
        cols[i] = ptr
    ctx._check(lib().kb_star_join_host_into(ctx.h, C.c_void_p(s_ptr), C.c_void_p(p_ptr), C.c_void_p(o_ptr), n, join_slot, patterns(pats), len(pats),
                                            a, nf, C.byref(n_cols), slots, cols, out_cap, C.byref(n_rows)))
    return n_rows.value, [slots[i] for i in range(n_cols.value)]


def make_rules(rules: Sequence[dict]):
    arr = (KbRule * max(len(rules), 1))()
    keep = []
    for i, r in enumerate(rules):
        prem = patterns(r["premise"])
        conc = patterns(r["conclusion"])
        fl = r.get("filters", [])
        farr = (KbRuleFilter * max(len(fl), 1))()
        for j, f in enumerate(fl):
            farr[j] = f
        keep += [prem, conc, farr]
        arr[i] = KbRule(C.cast(prem, C.POINTER(KbPattern)), len(r["premise"]), C.cast(farr, C.POINTER(KbRuleFilter)), len(fl),
                        C.cast(conc, C.POINTER(KbPattern)), len(r["conclusion"]))
    arr._keep = keep  # keep the nested arrays alive as long as the rule array
    return arr, keep


def legacy_hash_join_cuda(subjects, predicates, objects, predicate_filter: int, literal_filter: Optional[int] = None, libpath: Optional[str] = None) -> np.ndarray:
    """Replays exactly what Kolibrie's `hash_join_cuda` does (kolibrie/src/cuda/cuda_join.rs:28-60) against the legacy symbol."""
    L = lib() if libpath is None else C.CDLL(libpath)
    fn = L.perform_hash_join_cuda
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint32)]
    s, p, o = _u32(subjects), _u32(predicates), _u32(objects)
    lit = C.c_uint32(literal_filter) if literal_filter is not None else None
    idx = C.POINTER(C.c_uint32)()
    cnt = C.c_uint32(0)
    fn(_ptr(s), _ptr(p), _ptr(o), len(s), predicate_filter, C.byref(lit) if lit is not None else None, C.byref(idx), C.byref(cnt))
    out = np.ctypeslib.as_array(idx, shape=(cnt.value,)).copy() if cnt.value else np.empty(0, np.uint32)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    if idx:
        libc.free(C.cast(idx, C.c_void_p))  # the Rust side frees it through Vec's drop = libc free
    return out


def segment_write(path: str, s, p, o, tag: int = 0):
    """kb_segment_write (host only): three u32 columns as a columnar segment file"""
    s, p, o = _u32(s), _u32(p), _u32(o)
    rc = lib().kb_segment_write(path.encode(), _ptr(s), _ptr(p), _ptr(o), len(s), tag)
    if rc != KB_OK:
        raise KolibrieError(rc, f"cannot write {path}")


def segment_info(path: str):
    """kb_segment_info (host only): (n_triples, tag, cmin[3], cmax[3]) of a segment file"""
    n, tag = C.c_uint64(), C.c_uint64()
    lo, hi = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
    rc = lib().kb_segment_info(path.encode(), C.byref(n), C.byref(tag), lo, hi)
    if rc != KB_OK:
        raise KolibrieError(rc, f"{path} is not a segment file")
    return n.value, tag.value, list(lo), list(hi)
