#!/bin/bash
# round 2, step ai: C2 (99 free keyframes) through the big-window path with the two-level preconditioner, A/B
set -x
for v in 129 64; do
  SQRTBA_BIG_MIN_SLOTS=$v python tools/single_window.py --config c2 --reps 3 2>&1 | grep "^{" | tail -1
done
SQRTBA_BIG_MIN_SLOTS=64 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "large_window_c2" 2>&1 | tail -3
