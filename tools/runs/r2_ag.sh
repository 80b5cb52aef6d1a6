#!/bin/bash
# round 2, step ag: full GPU suite after the preconditioner work + C3 timing + per-launch durations
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/gba_sharded.py --pcg-mode 0 2>&1 | grep "^{" | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_chunk|k_coarse|k_cg_prep" -c 12 --csv --log-file gpurun_out/r2_ag_launches.csv python tools/gba_sharded.py --reps 1 > gpurun_out/r2_ag_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_ag_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:13]: print(r[ki][:40], r[vi], r[ui])
P
