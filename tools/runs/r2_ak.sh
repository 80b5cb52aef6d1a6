#!/bin/bash
# round 2, step ak: automatic PCG tolerance (1e-7 on local windows): full GPU suite, short bench, single-window timings
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python tools/single_window.py --config c0 --reps 5 2>&1 | grep "^{" | tail -1
python tools/single_window.py --config c2 --reps 3 2>&1 | grep "^{" | tail -1
python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/r2_ak_bench.json 2> gpurun_out/r2_ak_bench.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2_ak_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['parity'], d['solve_stats_last_step']['cg_iters_total'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
P
