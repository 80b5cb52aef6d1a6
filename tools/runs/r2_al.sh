#!/bin/bash
# round 2, step al: host-side timeline of a 256-window sqrtba_set_problem_batch (SQRTBA_HOST_TIMING) + GBA tolerance probe
set -x
nproc
SQRTBA_HOST_TIMING=1 python bench.py --steps 2 --warmup 1 --skip-extras --skip-cpu-baseline 2>&1 | grep "sqrtba host" | tail -40
python tools/rtol_probe.py 2>&1 | grep "^{" | head -3
