#!/bin/bash
# round 2, step x: chunk preconditioner (20-pose blocks) for the big-window persistent PCG -- parity, then A/B against
# the 6x6 block-Jacobi blocks (pcg_mode 5) on C3 (one GPU) and on one shard of the 8-way split
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "big_window or global_ba_loop or chunk" 2>&1 | tail -15
for m in 0 5; do
  python tools/gba_sharded.py --pcg-mode $m 2>&1 | grep "^{" | tail -1
  python tools/gba_proxy.py --nshards 8 --pcg-mode $m 2>&1 | grep "^{" | tail -1
done
