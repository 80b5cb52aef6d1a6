#!/bin/bash
# round 2, step ae: ncu --set full of the chunk-factor and coarse-inversion kernels (where do 230 / 490 us go)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_chunk_factor|k_coarse_invert" -s 2 -c 2 -o gpurun_out/r2_ae_prec -f python tools/gba_sharded.py --reps 1 > gpurun_out/r2_ae_ncu.log 2>&1
ls -la gpurun_out/r2_ae_prec.ncu-rep
