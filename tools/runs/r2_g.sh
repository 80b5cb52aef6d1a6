#!/bin/bash
# where the end-to-end overhead of a 256-window batch goes (host steps of set_problem, uploads, read-back)
SQRTBA_HOST_TIMING=1 python - <<'PY' 2>&1 | tail -60
import sys, time
sys.path.insert(0,'.')
from bench import load_pkg, make_batch, pinned_copy
import numpy as np
pkg=load_pkg()
wins, prob, pp, tp, op = make_batch(pkg, 256, 0)
host = [pinned_copy(a) for a in (prob.pose_qt, prob.pose_fixed, prob.cam, prob.point_xyz, prob.obs_pose, prob.obs_point, prob.obs_meas)]
hprob = pkg.synth.Problem(*[t.numpy() for t in host])
ba = pkg.SqrtBA(host_threads=16)
for rep in range(3):
    print('--- rep', rep, file=sys.stderr)
    t0=time.perf_counter(); ba.set_problem_batch(hprob, pp, tp, op); t1=time.perf_counter()
    st=ba.solve_local(); t2=time.perf_counter()
    ba.poses(); ba.points(); ba.outliers(); t3=time.perf_counter()
    print(f'set_problem {1e3*(t1-t0):.1f} ms, solve {1e3*(t2-t1):.1f} ms (device {st["ms_total"]:.1f}), read-back {1e3*(t3-t2):.1f} ms', file=sys.stderr)
PY
