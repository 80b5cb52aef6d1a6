"""PCG tolerance against the result tolerances of north_star on local-BA windows (C0 shape, several seeds): identical
trial sequence and outlier flags, cost 1e-6, pose RMS 1e-5 m / 1e-6 rad -- how far is pcg_rtol from breaking them?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, sys.path[0] + "/tests")
import numpy as np  # noqa: E402
from bench import load_pkg  # noqa: E402

pkg = load_pkg()
from oracle import refba  # noqa: E402
from test_gpu_parity import pose_rms  # noqa: E402

seeds = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3,4,5".split(","))]
rtols = tuple(float(x) for x in os.environ.get("RTOLS", "1e-9,3e-9,1e-8,3e-8,1e-7").split(","))
refs = {}
for seed in seeds:
    prob = pkg.synth.config_c0(seed)
    ref = refba.RefBA(prob, threads=8)
    ref.solve_local(0)
    refs[seed] = (prob, ref.trace(), ref.poses(), ref.outliers())
for rtol in rtols:
    worst = dict(rtol=rtol, same=True, flags=True, cost_rel=0.0, t_rms=0.0, r_rms=0.0, cg=0)
    ba = pkg.SqrtBA(pcg_rtol=rtol)
    for seed in seeds:
        prob, tr, pr, fr = refs[seed]
        ba.set_problem(prob)
        st = ba.solve_local()
        tg = ba.trace()
        same = len(tg) == len(tr) and np.array_equal(tg[:, [0, 1, 2, 7]], tr[:, [0, 1, 2, 7]])
        worst["same"] &= bool(same)
        worst["flags"] &= bool(np.array_equal(ba.outliers(), fr))
        if same:
            worst["cost_rel"] = max(worst["cost_rel"], float(np.max(np.abs(tg[:, 5] - tr[:, 5]) / tr[:, 5])))
        t, r = pose_rms(ba.poses(), pr, prob.pose_fixed == 0)
        worst["t_rms"], worst["r_rms"] = max(worst["t_rms"], t), max(worst["r_rms"], r)
        worst["cg"] += st["cg_iters_total"]
    ba.close()
    print(json.dumps(worst), flush=True)
