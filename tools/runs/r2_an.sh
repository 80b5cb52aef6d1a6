#!/bin/bash
# round 2, step an (2 GPUs): preconditioner levels refreshed every 2nd LM iteration -- sharded parity (2 ranks), single-GPU
# global-BA parity, C3 at 2 and 1 GPUs
set -x
timeout 500 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "global_ba or big_window or chunk" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gba_sharded.py --skip-single 2>&1 | grep "^{" | tail -1
timeout 300 python tools/gba_sharded.py 2>&1 | grep "^{" | tail -1
