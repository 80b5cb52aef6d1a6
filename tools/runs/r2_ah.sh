#!/bin/bash
# round 2, step ah (8 GPUs): sharded global BA with the two-level preconditioner -- parity against the oracle (4 ranks),
# C3 time-to-converge at 8 / 4 / 2 GPUs, chunk-only and 6x6 variants at 8 GPUs for the A/B
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 tools/gba_sharded.py --skip-single --pcg-mode $2 2>&1 | grep "^{" | tail -1; }
run 8 0
run 8 6
run 4 0
run 2 0
