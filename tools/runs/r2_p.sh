#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
from bench import load_pkg, make_batch
pkg=load_pkg()
wins, prob, pp, tp, op = make_batch(pkg, 64, 0)
for mode in (0, 4):
    ba=pkg.SqrtBA(pcg_mode=mode); ba.set_problem_batch(prob,pp,tp,op)
    ms=[]
    for _ in range(3):
        ba.reset_state(); st=ba.solve_local(); ms.append(st["ms_total"])
    print('64 windows pcg_mode',mode,'ms',min(ms), 'launches', st['kernel_launches'], 'reproducible', st['reproducible'])
    ba.close()
w=pkg.synth.config_c0(0)
for mode in (0,4):
    ba=pkg.SqrtBA(pcg_mode=mode); ba.set_problem(w); ms=[]
    for _ in range(4):
        ba.reset_state(); st=ba.solve_local(); ms.append(st["ms_total"])
    print('C0 pcg_mode',mode,'ms',min(ms)); ba.close()
PY
