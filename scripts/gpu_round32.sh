#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== decode throughput"; timeout 300 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c
ctx = c.Context(0)
n_ids, n_rows = 2_000_000, 16_666_667
strings = ["http://example.org/employee%d" % i for i in range(n_ids)]
ctx.dict_strings_load(strings)
rel = ctx.rel_from_host([0], [np.random.default_rng(1).integers(0, n_ids, n_rows).astype(np.uint32)])
import ctypes as C
L = c.lib()
for rep in range(3):
    ctx.synchronize(); t0 = time.perf_counter()
    h = C.c_void_p(); ctx._check(L.kb_rel_decode(ctx.h, rel.h, 0, C.byref(h))); ctx.synchronize()
    dt = time.perf_counter() - t0
    n, tot = C.c_uint64(), C.c_uint64(); L.kb_strings_info(h, C.byref(n), C.byref(tot)); L.kb_strings_free(ctx.h, h)
print("decode %d rows -> %d bytes in %.3f ms (%.1f GB/s of output)" % (n.value, tot.value, dt * 1e3, tot.value / dt / 1e9))
PY
