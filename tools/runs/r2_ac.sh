#!/bin/bash
# round 2, step ac: two-level preconditioner with warp-reduced cross-chunk sums: parity, C3 timing, per-launch durations
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "big_window or global_ba or chunk" 2>&1 | grep -v "^$" | tail -4
timeout 300 python tools/gba_sharded.py --pcg-mode 0 2>&1 | grep "^{" | tail -1
timeout 300 python tools/gba_proxy.py --nshards 8 --pcg-mode 0 2>&1 | grep "^{" | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_chunk|k_coarse|k_pcg_persist|k_cg_prep" -c 24 --csv --log-file gpurun_out/r2_ac_launches.csv python tools/gba_sharded.py --reps 1 > gpurun_out/r2_ac_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_ac_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:25]: print(r[ki][:40], r[vi], r[ui])
P
