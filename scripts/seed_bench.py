"""Single GPU: incremental materialisation (kb_datalog_fixpoint_seed) against starting over, on the config-4 shape at 1/10 scale
(111 111 classes, 4.9 M rdf:type facts): the closure of everything but the last 1 % of the instances, then that 1 % as five seeds of
0.2 % each; beside it one closure of all the facts from scratch. Checks that both ways end with the same number of facts."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, datagen

t = datagen.taxonomy_dataset(fanout=10, depth=5, n_instances=4_888_889)
rules = datagen.taxonomy_rules(t)
n = len(t.s)
late = n - 48_000
res = {"triples": int(n), "seed_triples_each": 9600}

cx = c.Context(0)
cx.store_load(t.s[:late], t.p[:late], t.o[:late])
cx.datalog_fixpoint(rules)[0].free()          # warm-up closure (pool growth, set sizing hints)
cx.store_load(t.s[:late], t.p[:late], t.o[:late])
t0 = time.perf_counter(); rel, st = cx.datalog_fixpoint(rules); cx.synchronize(); w = time.perf_counter() - t0
inferred = int(st.inferred); rel.free()
res["closure_of_99pct"] = {"wall_ms": w * 1e3, "device_ms": st.device_ms, "inferred": inferred}
seeds = []
for k in range(5):
    a, b = late + k * 9600, late + (k + 1) * 9600
    seed = cx.rel_from_host([0, 1, 2], [t.s[a:b].copy(), t.p[a:b].copy(), t.o[a:b].copy()])
    cx.synchronize()
    t0 = time.perf_counter(); out, n_new, st2 = cx.datalog_fixpoint_seed(rules, seed); cx.synchronize(); w = time.perf_counter() - t0
    assert n_new == 9600
    inferred += int(st2.inferred)
    seeds.append({"wall_ms": w * 1e3, "device_ms": st2.device_ms, "accepted": int(n_new), "inferred": int(st2.inferred), "rounds": int(st2.rounds)})
    out.free(); seed.free()
res["seeds"] = seeds
cx.close()

cx = c.Context(0)
cx.store_load(t.s, t.p, t.o)
cx.datalog_fixpoint(rules)[0].free()
cx.store_load(t.s, t.p, t.o)
t0 = time.perf_counter(); rel, st = cx.datalog_fixpoint(rules); cx.synchronize(); w = time.perf_counter() - t0
res["closure_from_scratch"] = {"wall_ms": w * 1e3, "device_ms": st.device_ms, "inferred": int(st.inferred)}
res["same_fact_count"] = bool(int(st.inferred) == inferred)
cx.close()
print(json.dumps(res))
assert res["same_fact_count"], (int(st.inferred), inferred)
