#!/bin/bash
python tools/stage_roofline.py --config c4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['kernels']['k_matvec'])"
python tools/stage_roofline.py --config c3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['kernels']['k_matvec'])"
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
