"""DRAM bytes per Datalog candidate: run under
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:derive_ --csv
one config-4 closure (second closure of the process: pool and known-fact sets warm); prints the derivation count the bytes are divided by."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolibrie_b200 import capi as c, datagen
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 48_888_890
t = datagen.taxonomy_dataset(10, 6, n_inst, seed=43)
rules = datagen.taxonomy_rules(t)
ctx = c.Context(0)
ctx.store_load(t.s, t.p, t.o)
rel, st = ctx.datalog_fixpoint(rules)
ctx.synchronize()
print("DERIVATIONS", int(st.derivations), "INFERRED", int(st.inferred), flush=True)
