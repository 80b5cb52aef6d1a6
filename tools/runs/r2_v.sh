#!/bin/bash
# round 2, step v: per-CTA cycle counts of the persistent PCG kernel (prof build) with tile features, to fit the static
# tile-range cost model: 8-way shard, 2-way shard, whole C3, C2
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_v_*.csv
export SQRTBA_LIB=sqrtlm-slam_b200/libsqrtba_prof.so
SQRTBA_PROF_CSV=gpurun_out/r2_v_shard8.csv python tools/gba_proxy.py --nshards 8 2>&1 | grep "persist prof\|^{" | tail -12
SQRTBA_PROF_CSV=gpurun_out/r2_v_shard1.csv python tools/gba_proxy.py --nshards 1 2>&1 | grep "persist prof\|^{" | tail -8
SQRTBA_PROF_CSV=gpurun_out/r2_v_c2.csv python tools/single_window.py --config c2 --reps 1 2>&1 | grep "persist prof\|^{" | tail -8
