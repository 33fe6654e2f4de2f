"""cfg4 twice in one process (cold pool, then warm) with KOLIBRIE_TRACE phase timings"""
import os, sys, time
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c, datagen
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 48_888_890
t = datagen.taxonomy_dataset(10, 6, n_inst, seed=43)
rules = datagen.taxonomy_rules(t)
ctx = c.Context(0)
for rep in range(3):
    ctx.store_load(t.s, t.p, t.o)
    ctx.synchronize()
    sys.stderr.write(f"==== run {rep}\n")
    t1 = time.time()
    rel, st = ctx.datalog_fixpoint(rules)
    ctx.synchronize()
    print(rep, "wall", round(time.time() - t1, 4), "device_ms", round(st.device_ms, 1), "inferred", st.inferred, flush=True)
    rel.free()
