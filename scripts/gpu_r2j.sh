#!/bin/bash
# round 2, step j: slide maintenance in one split pass, buffer recycling
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_gpu_rsp.py tests/test_gpu_plan.py tests/test_rdf_star_rsp_golden.py tests/test_gpu_parity.py tests/test_gpu_datalog.py -m gpu -x -q 2>&1 | tail -15
echo "tests: $((SECONDS-T0)) s"
timeout 300 python bench.py --config cfg5 > gpurun_out/cfg5_r2j.json 2> gpurun_out/cfg5_r2j.err; tail -c 1500 gpurun_out/cfg5_r2j.json; tail -3 gpurun_out/cfg5_r2j.err
KOLIBRIE_INDEX_SPLIT=0 timeout 300 python bench.py --config cfg5 > gpurun_out/cfg5_r2j_nosplit.json 2>/dev/null; tail -c 700 gpurun_out/cfg5_r2j_nosplit.json
timeout 300 python bench.py --config cfg4 > gpurun_out/cfg4_r2j.json 2>/dev/null; tail -c 900 gpurun_out/cfg4_r2j.json
T1=$SECONDS
timeout 600 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench: $((SECONDS-T1)) s"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2j.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','e2e','roofline','gpu_launches') if k in l})
d=l.get('details',{})
for k in ('sync_path','scan_path','cfg2_10M','adversarial'):
    if k in d: print(k, json.dumps(d[k])[:600])
for k,v in d.get('other_configs',{}).items(): print(k, json.dumps(v)[:500])
PY
