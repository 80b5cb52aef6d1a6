"""PCG tolerance on the full C3 global BA (two-level preconditioner) against the oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, sys.path[0] + "/tests")
import numpy as np
from bench import load_pkg
pkg = load_pkg()
from oracle import refba
from test_gpu_parity import pose_rms
prob = pkg.synth.config_c3(0)
t0 = time.time()
ref = refba.RefBA(prob, threads=max(1, min(os.cpu_count() or 1, 32))); ref.solve_global(10, False)
print("oracle s", time.time() - t0, flush=True)
tr = ref.trace()
for rtol in tuple(float(x) for x in os.environ.get("RTOLS", "1e-9,1e-8,1e-7,1e-6").split(",")):
    ba = pkg.SqrtBA(pcg_rtol=rtol)
    ba.set_problem(prob)
    st = ba.solve_global(10, False)
    tg = ba.trace()
    t, r = pose_rms(ba.poses(), ref.poses(), prob.pose_fixed == 0)
    same = len(tg) == len(tr) and np.array_equal(tg[:, [0, 1, 2, 7]], tr[:, [0, 1, 2, 7]])
    print(json.dumps(dict(rtol=rtol, t_rms=t, r_rms=r, cg=st["cg_iters_total"], same=bool(same), ms=st["ms_total"],
          cost_rel=float(np.max(np.abs(tg[:, 5] - tr[:, 5]) / tr[:, 5])) if same else None,
          pts_max=float(np.abs(ba.points() - ref.points()).max()))), flush=True)
    ba.close()
