#!/bin/bash
# round 2, record run after the big-window threshold change (64 free keyframes): full GPU suite, smoke, bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_ar_pytest.log 2>&1; tail -4 gpurun_out/r2_ar_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ar_bench.json 2> gpurun_out/r2_ar_bench.err; tail -c 300 gpurun_out/r2_ar_bench.err; wc -c gpurun_out/r2_ar_bench.json
