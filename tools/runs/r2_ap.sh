#!/bin/bash
# round 2, final record run: full GPU suite, bench (both arms), ncu launch list of the bench command, ncu --set full of
# the dominant kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_ap_pytest.log 2>&1; tail -4 gpurun_out/r2_ap_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ap_bench.json 2> gpurun_out/r2_ap_bench.err; tail -c 300 gpurun_out/r2_ap_bench.err; wc -c gpurun_out/r2_ap_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ap_bench_ref.json 2> gpurun_out/r2_ap_bench_ref.err; tail -c 400 gpurun_out/r2_ap_bench_ref.json; echo
python bench.py --steps 1 --warmup 3 --windows-per-gpu 64 --skip-extras --skip-cpu-baseline > gpurun_out/r2_ap_small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_ap_launches_64win.csv python bench.py --steps 1 --warmup 3 --windows-per-gpu 64 --skip-extras --skip-cpu-baseline > gpurun_out/r2_ap_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ap_ncu_launches.log | cut -c1-200
python tools/prof_stage.py 64 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_matvec_pipe -s 2 -c 1 -o gpurun_out/r2_ap_matvec -f python tools/prof_stage.py 64 3 > gpurun_out/r2_ap_ncu_full.log 2>&1
ls -la gpurun_out/r2_ap_matvec.ncu-rep
