#!/bin/bash
# final single-GPU pass of the round: every GPU test, smoke, the reference arm, the default bench line
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "gpu tests: $((SECONDS-T0)) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T1=$SECONDS
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r2z.json 2>/dev/null; echo "reference arm: $((SECONDS-T1)) s"; tail -c 300 gpurun_out/bench_ref_r2z.json
T2=$SECONDS
timeout 600 python bench.py > gpurun_out/bench_full_r2z.json 2> gpurun_out/bench_full_r2z.err; echo "bench: $((SECONDS-T2)) s"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_full_r2z.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks') if k in l})
print('e2e', l['e2e']['value'], 'roofline', l['roofline']['frac'], 'traffic', l['roofline'].get('traffic'))
print('cpu_baseline', l['cpu_baseline'])
for k in ('sync_path','scan_path','cfg2_10M','adversarial'):
    if k in l: print(k, json.dumps(l[k])[:500])
for k,v in l.get('other_configs',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('workload','roofline','parity','cpu_baseline')})[:600])
PY
