#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_h.log 2>&1; tail -4 gpurun_out/r2_pytest_h.log
python tools/single_window.py --config c0; python tools/single_window.py --config c1; python tools/single_window.py --config c2
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null > gpurun_out/r2_bench_h.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_h.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(json.dumps(d['global_ba'])[:330])"
