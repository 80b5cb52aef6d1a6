#!/bin/bash
# matvec residency A/B after the 104-byte columns
python tools/stage_roofline.py --config c4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 auto', d['kernels'])"
python tools/stage_roofline.py --config c3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 auto', d['kernels'])"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from bench import load_pkg, make_batch
pkg=load_pkg()
_, prob, pp, tp, op = make_batch(pkg, 256, 0)
for S in (2,3):
    ba=pkg.SqrtBA(pipe_stages=S); ba.set_problem_batch(prob,pp,tp,op); ba.debug_linearize(1); ba.debug_step(100.0)
    print('c4 stages',S,'matvec ms', ba.time_stage(0,3,20)); ba.close()
PY
bash tools/runs/r2_c.sh
