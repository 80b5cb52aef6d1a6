#!/bin/bash
# round 2, step ao (8 GPUs): final C3 time-to-converge at 8 and 4 GPUs (two-level preconditioner refreshed every 2nd LM
# iteration, PCG tolerance 1e-8)
set -x
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 tools/gba_sharded.py --skip-single --pcg-mode $2 2>&1 | grep "^{" | tail -1; }
run 8 0
run 4 0
