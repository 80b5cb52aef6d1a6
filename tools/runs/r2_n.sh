#!/bin/bash
# round-2 record run: full GPU suite, the bench line (both arms), the ncu launch list of the bench command
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_n.log 2>&1; tail -4 gpurun_out/r2_pytest_n.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n.json 2> gpurun_out/r2_bench_ref_n.err; tail -c 300 gpurun_out/r2_bench_ref_n.json; echo
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n.json 2> gpurun_out/r2_bench_n.err; tail -c 300 gpurun_out/r2_bench_n.err
python bench.py --steps 1 --warmup 3 --windows-per-gpu 64 --skip-extras --skip-cpu-baseline > gpurun_out/r2_bench_small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_64win.csv python bench.py --steps 1 --warmup 3 --windows-per-gpu 64 --skip-extras --skip-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-300
