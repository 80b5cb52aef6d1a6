#!/bin/bash
python tools/prof_stage.py 64 3 > gpurun_out/r2_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_matvec_pipe -s 2 -c 1 -o gpurun_out/r02_matvec_compactJ python tools/prof_stage.py 64 3 > gpurun_out/r2_ncu_f.log 2>&1
tail -3 gpurun_out/r2_ncu_f.log
