"""config 4 shape on one GPU: 10-ary class tree of depth 6 + N rdf:type facts, rules R1/R2; checks closed-form counts."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from kolibrie_b200 import capi as c, datagen

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
fan, depth = 10, 6
t0 = time.time()
t = datagen.taxonomy_dataset(fan, depth, n_inst)
rules = datagen.taxonomy_rules(t)
ctx = c.Context(0)
ctx.store_load(t.s, t.p, t.o)
ctx.set_timing(True)
t1 = time.time()
rel, st = ctx.datalog_fixpoint(rules)
t2 = time.time()
# closed form
cls = t.o[t.p == t.ids["rdf:type"]].astype(np.int64) - 2
lvl = np.zeros(t.n_classes, dtype=np.int64)
start = 0
for k in range(depth + 1):
    lvl[start:start + fan ** k] = k
    start += fan ** k
want_type = int(lvl[cls].sum())
want_sc = sum(fan ** k * (k - 1) for k in range(2, depth + 1))
rows = None
n_type = n_sc = -1
if st.inferred < 400_000_000:
    pcol = rel.column(1)
    n_type = int((pcol == t.ids["rdf:type"]).sum()); n_sc = int((pcol == t.ids["rdfs:subClassOf"]).sum())
stats = ctx.get_stats()
print(json.dumps({"triples": len(t.s), "inferred": int(st.inferred), "rounds": int(st.rounds), "round_new": [int(st.round_new[i]) for i in range(st.rounds)],
                  "derivations": int(st.derivations), "device_ms": st.device_ms, "wall_s": t2 - t1, "gen_s": t1 - t0,
                  "type_ok": n_type == want_type, "sc_ok": n_sc == want_sc, "want_type": want_type, "want_sc": want_sc,
                  "facts_per_s": st.inferred / (t2 - t1), "stats": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in stats.items()}}))
