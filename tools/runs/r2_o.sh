#!/bin/bash
python -m pytest tests/test_gpu_adapter.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
python bench.py --steps 1 --warmup 3 --windows-per-gpu 16 2>gpurun_out/r2_bench_o.err > gpurun_out/r2_bench_o.json; tail -c 400 gpurun_out/r2_bench_o.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_o.json')); print(json.dumps(d['pose_only'])); print(json.dumps(d['essential_graph']))"
