#!/bin/bash
# round 2, step y: chunk preconditioner with FP32 vector reductions + division-free factor kernel: parity, C3 A/B,
# per-launch durations of the new kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "big_window or global_ba or chunk" 2>&1 | tail -5
for m in 0 5; do
  python tools/gba_sharded.py --pcg-mode $m 2>&1 | grep "^{" | tail -1
  python tools/gba_proxy.py --nshards 8 --pcg-mode $m 2>&1 | grep "^{" | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_chunk -c 12 --csv --log-file gpurun_out/r2_y_chunk_launches.csv python tools/gba_sharded.py --reps 0 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2_y_chunk_launches.csv | tail -12
