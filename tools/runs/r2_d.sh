#!/bin/bash
# compressed pose Jacobian (4 geometry rows instead of 18): parity + stage rates + latency
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_d.log 2>&1; tail -15 gpurun_out/r2_pytest_d.log
python tools/stage_roofline.py 2>&1 | tail -12
python tools/single_window.py --config c0; python tools/single_window.py --config c2
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null > gpurun_out/r2_bench_d.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_d.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(json.dumps(d['roofline'])[:900]); print(json.dumps(d['global_ba'])[:400])"
