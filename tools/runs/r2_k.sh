#!/bin/bash
# multi-GPU: sharded global BA against the oracle (pytest, 4 ranks when >= 4 GPUs), then C3 time-to-converge at 8 and 2 GPUs
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5
for n in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n tools/gba_sharded.py --skip-single 2>&1 | grep "^{" | tail -1
done
python tools/gba_sharded.py 2>&1 | grep "^{" | tail -1
