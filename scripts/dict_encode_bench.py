"""kb_dict_encode throughput: the employee shape's term strings (3 per triple, in document order), one call per 3 M-term batch;
the sequential Dictionary::encode loop of the Python mirror beside it on a sample."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, engine as E

n_emp = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
terms = []
for e in range(n_emp):
    s = f"http://example.org/employee{e}"
    for p, o in (("http://xmlns.com/foaf/0.1/name", s), ("http://xmlns.com/foaf/0.1/title", "Developer" if e % 3 else "Manager"),
                 ("https://data.cityofchicago.org/resource/xzkq-xp2w/annual_salary", str(50000 + e * 7919 % 120000))):
        terms += [s, p, o]
n = len(terms)
enc = [t.encode() for t in terms]
off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum([len(b) for b in enc], dtype=np.uint64)
data = np.frombuffer(b"".join(enc), dtype=np.uint8).copy()
ctx = c.Context(0)
lib = c.lib()
import ctypes as C
ids = np.empty(n, np.uint32); first = np.empty(n, np.uint64); n_new = C.c_uint32(0)
B = 3_000_000
for rep in range(2):
    ctx.dict_strings_load([])
    ctx.synchronize()
    t0 = time.perf_counter()
    for a in range(0, n, B):
        b = min(n, a + B)
        o = (off[a:b + 1] - off[a]).copy()
        rc = lib.kb_dict_encode(ctx.h, o.ctypes.data, data[int(off[a]):].ctypes.data, b - a, ids[a:].ctypes.data, C.byref(n_new), first.ctypes.data)
        assert rc == 0
    ctx.synchronize()
    dt = time.perf_counter() - t0
print(f"kb_dict_encode: {n} terms ({len(data) / 1e6:.0f} MB of strings), batches of {B}: {dt * 1e3:.1f} ms = {n / dt / 1e6:.1f} M terms/s "
      f"(host->device copy of the strings and device->host copy of the ids included); ids on the device dictionary: {ctx.dict_strings_info()[0]}")
m = min(n, 3_000_000)
d = E.Dictionary()
t0 = time.perf_counter()
want = np.array([d.encode(t) for t in terms[:m]], dtype=np.uint32)
dt_py = time.perf_counter() - t0
assert np.array_equal(want, ids[:m])
print(f"sequential encode loop of the Python mirror on the first {m} terms: {dt_py * 1e3:.0f} ms = {m / dt_py / 1e6:.1f} M terms/s (ids identical)")
