#!/bin/bash
# round 2: ncu launch list of one global BA (C3, one GPU) on the final tree -- per-kernel shares
set -x
python tools/gba_sharded.py --reps 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_at_gba_launches.csv python tools/gba_sharded.py --reps 1 > gpurun_out/r2_at_ncu.log 2>&1
python - <<'P'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/r2_at_gba_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    name=re.sub(r'\(.*','',r[ki]).replace('void ','')
    v=float(r[vi].replace(',','')); us=v/1000 if r[ui] in ('ns','nsecond') else v
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=us
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k},{n},{t:.1f},{t/n:.2f},{t/tot:.4f}")
P
