import sys, json
sys.path.insert(0,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, sys.path[0] + '/tests')
import numpy as np
from bench import load_pkg
pkg = load_pkg()
from oracle import refba
from test_gpu_parity import pose_rms
prob = pkg.synth.make_problem(17, 200, 1, 8000, 9.0, stereo=True, loop=True, cand_halfwidth=15)
for robust in (False, True):
    ref = refba.RefBA(prob); ref.solve_global(10, robust)
    for rtol in tuple(float(x) for x in __import__("os").environ.get("RTOLS", "1e-9,1e-10,1e-11,1e-12,1e-13").split(",")):
        ba = pkg.SqrtBA(pcg_rtol=rtol, pcg_max_iters=2000)
        ba.set_problem(prob)
        st = ba.solve_global(10, robust)
        tg, tr = ba.trace(), ref.trace()
        t, r = pose_rms(ba.poses(), ref.poses(), prob.pose_fixed == 0)
        same = len(tg)==len(tr) and np.array_equal(tg[:, [0,1,2,7]], tr[:, [0,1,2,7]])
        print(json.dumps(dict(robust=robust, rtol=rtol, t_rms=t, r_rms=r, cg=st["cg_iters_total"], same=bool(same),
              cost_rel=float(np.max(np.abs(tg[:,5]-tr[:,5])/tr[:,5])) if same else None, ms=st["ms_total"])), flush=True)
        ba.close()
