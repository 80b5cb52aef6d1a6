#!/bin/bash
# round 2, step w: tile ranges of the persistent PCG kernel balanced by the fitted cost model -- proxy shard, C2, C0, C3
set -x
mkdir -p gpurun_out
SQRTBA_LIB=sqrtlm-slam_b200/libsqrtba_prof.so python tools/gba_proxy.py --nshards 8 2>&1 | grep "persist prof\] grid\|slowest\|fastest\|^{" | tail -4
python tools/gba_proxy.py --nshards 8 2>&1 | grep "^{" | tail -1
python tools/gba_proxy.py --nshards 1 2>&1 | grep "^{" | tail -1
python tools/single_window.py --config c2 --reps 3 2>&1 | grep "^{" | tail -1
python tools/single_window.py --config c0 --reps 5 2>&1 | grep "^{" | tail -1
python tools/single_window.py --config c1 --reps 5 2>&1 | grep "^{" | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
