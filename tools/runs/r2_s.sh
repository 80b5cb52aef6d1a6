#!/bin/bash
# round 2, step s: OptimizeSim3 on the device -- new GPU tests, then the whole GPU suite and the bench on the final tree
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sim3_opt.py tests/test_gpu_adapter.py -x -q > gpurun_out/r2_s_sim3.log 2>&1; tail -15 gpurun_out/r2_s_sim3.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_s_pytest.log 2>&1; tail -8 gpurun_out/r2_s_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_s_bench.json 2> gpurun_out/r2_s_bench.err; tail -c 400 gpurun_out/r2_s_bench.err; wc -c gpurun_out/r2_s_bench.json
