#!/bin/bash
mkdir -p gpurun_out
python scripts/cfg5_breakdown.py 2>&1 | tail -8
KOLIBRIE_TRACE=1 python scripts/cfg5_breakdown.py 2>&1 | grep "kb trace" | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"segment_|clear_chunks|probe_index|build_direct" --csv --log-file gpurun_out/cfg5_kernels.csv python scripts/cfg5_breakdown.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/cfg5_kernels.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
acc=collections.defaultdict(list)
for r in rows[1:]:
    acc[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1))
for k,v in acc.items(): print(k, len(v), 'launches, median us', sorted(v)[len(v)//2], 'last', v[-1])
PY
