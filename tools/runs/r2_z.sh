#!/bin/bash
# round 2, step z: per-launch durations of the chunk-preconditioner kernels on C3 (ncu launch list)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_chunk|k_pcg_persist|k_qr_pipe2|k_linearize" -c 40 --csv --log-file gpurun_out/r2_z_chunk_launches.csv python tools/gba_sharded.py --reps 1 > gpurun_out/r2_z_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_z_chunk_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:41]: print(r[ki][:40], r[vi], r[ui])
P
