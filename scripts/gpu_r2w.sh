#!/bin/bash
cat > /tmp/dl8.py <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import torch
from kolibrie_b200 import capi as c, datagen
t = datagen.taxonomy_dataset(10, 6, 48_888_890, seed=43)
rules = datagen.taxonomy_rules(t)
ctx = c.Context(0)
def used():
    f, tot = torch.cuda.mem_get_info(0)
    return (tot - f) / 2**30
ws = []
for rep in range(16):
    ctx.store_load(t.s, t.p, t.o)
    ctx.synchronize()
    t1 = time.time()
    rel, st = ctx.datalog_fixpoint(rules, c.SEMI_NAIVE if rep < 12 else c.SEMI_NAIVE_OLD_DELTA)
    ctx.synchronize()
    ws.append(round(time.time() - t1, 4))
    rel.free()
print("walls", ws, "memory in use", round(used(), 1), "GiB", flush=True)
PY
python /tmp/dl8.py 2>&1 | tail -2
