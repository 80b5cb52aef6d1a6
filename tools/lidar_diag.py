"""Diagnostic: GPU vs oracle traces of the lidar pass on one problem (prints pass-3 rows side by side)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sqrtlm-slam_b200")
from oracle import refba
synth = pkg.synth
np.set_printoptions(linewidth=200, precision=10)
prob = synth.make_problem(33, 12, 4, 700, 6.0, stereo=True, name="adapter-lidar")
ld = synth.lidar_data(prob, seed=3, n_flat=500, n_corner=120)
for numeric in (True, False):
    h = pkg.SqrtBA(third_pass_iters=20)
    h.set_problem(prob); h.set_lidar(ld, numeric_jacobian=numeric); h.solve_local()
    r = refba.RefBA(prob); r.set_lidar(ld, numeric_jacobian=numeric); r.solve_local(20)
    tg, tr = h.trace(), r.trace()
    print("numeric", numeric, "rows", len(tg), len(tr), "edges", h.num_lidar_edges(), r.num_lidar_edges(),
          "match diff", int((h.lidar_matches() != r.lidar_matches()).sum()))
    n = min(len(tg), len(tr))
    for i in range(n):
        if tg[i, 0] == 2 or tr[i, 0] == 2:
            print("G", tg[i, :8]); print("O", tr[i, :8])
    for i in range(n, len(tg)): print("G+", tg[i, :8])
    for i in range(n, len(tr)): print("O+", tr[i, :8])
    d = np.abs(h.poses() - r.poses())
    print("max pose diff t %.3e q %.3e" % (d[:, :3].max(), d[:, 3:].max()))
    h.close()
