#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:derive_ --csv --log-file gpurun_out/derive_traffic_r2s.csv python scripts/derive_traffic.py > gpurun_out/derive_traffic_r2s.txt 2>&1
tail -2 gpurun_out/derive_traffic_r2s.txt
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/derive_traffic_r2s.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
acc=collections.defaultdict(lambda: collections.defaultdict(float)); n=collections.Counter()
mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-3,'us':1,'ms':1e3,'usecond':1,'nsecond':1e-3,'msecond':1e3}
for r in rows[1:]:
    k=r[ki].split('(')[0]; acc[k][r[mi]]+=float(r[vi].replace(',',''))*mult.get(r[ui],1)
    if r[mi].startswith('gpu__time'): n[k]+=1
for k,v in acc.items(): print(k, n[k], 'launches', {a:round(b/1e6,1) for a,b in v.items()}, '(MB / us*1e-6)')
PY
