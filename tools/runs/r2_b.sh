#!/bin/bash
# round-2 check B: GPU tests, single-window latency with / without the macro-step graph, GBA proxy after the owner-phase changes
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1; tail -15 gpurun_out/r2_pytest_b.log
for m in 0 3; do python tools/single_window.py --config c0 --pcg-mode $m; done
for m in 0 3; do python tools/single_window.py --config c2 --pcg-mode $m; done
SQRTBA_HOST_TIMING=1 python tools/single_window.py --config c0 --reps 1 2>&1 | grep -i "graph" | head -5
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --windows-per-gpu 32 2>/dev/null > gpurun_out/r2_bench_b.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b.json')); print(json.dumps(d['single_window'])); print(json.dumps(d['global_ba'])[:700])"
python tools/gba_proxy.py --nshards 8 2>&1 | tail -1
