"""Timing of the lidar tight-coupling pass (row N4) on the KITTI-00-shaped window C0: sqrtba_solve_local with
third_pass_iters = 20, with and without lidar clouds, next to the oracle on the host.  Prints one JSON line."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sqrtlm-slam_b200")
from oracle import refba  # noqa: E402

prob = pkg.synth.config_c0(0)
ld = pkg.synth.lidar_data(prob, seed=0)
out = {"workload": "C0 + lidar: %d flat + %d corner features, local map %d + %d points" %
       (len(ld.flat_xyz), len(ld.corner_xyz), len(ld.map_flat_xyz), len(ld.map_corner_xyz))}
h = pkg.SqrtBA(third_pass_iters=20)
for name, with_lidar in (("plain_third_pass", False), ("lidar_pass", True)):
    ms = []
    for rep in range(5):
        h.set_problem(prob)
        if with_lidar:
            h.set_lidar(ld)
        st = h.solve_local()
        ms.append(st["ms_total"])
    out[name + "_ms"] = float(np.median(ms[1:]))
    out[name + "_launches"] = st["kernel_launches"]
    tr = h.trace()
    out[name + "_trials_pass3"] = int((tr[:, 0] == 2).sum())
out["lidar_edges"] = h.num_lidar_edges()
r = refba.RefBA(prob)
r.set_lidar(ld)
t = time.time()
r.solve_local(20)
out["oracle_cpu_s"] = time.time() - t
out["oracle_lidar_edges"] = r.num_lidar_edges()
print(json.dumps(out))
