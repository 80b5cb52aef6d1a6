#!/bin/bash
# round 2, step ab: two-level preconditioner (chunk blocks + coarse correction over the chunks): parity, C3 A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "big_window or global_ba or chunk" 2>&1 | grep -v "^$" | tail -8
for m in 0 6; do
  timeout 300 python tools/gba_sharded.py --pcg-mode $m 2>&1 | grep "^{" | tail -1
  timeout 300 python tools/gba_proxy.py --nshards 8 --pcg-mode $m 2>&1 | grep "^{" | tail -1
done
