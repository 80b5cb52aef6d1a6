#!/bin/bash
# per-kernel durations of ONE C0 window (two-pass local BA), no graph so that every launch is listed in stream order
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_c0_launches.csv python tools/single_window.py --config c0 --reps 1 --pcg-mode 3 > gpurun_out/r2_c0_ncu.log 2>&1
tail -2 gpurun_out/r2_c0_ncu.log
./tools/microbench/gridbar
