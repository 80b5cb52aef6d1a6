#!/usr/bin/env python
"""Small driver for ncu: builds a batch of C0-shaped windows, runs one linearise + QR so that the planes hold real
data, then launches the requested stage kernels a few times through sqrtba_time_stage.
usage: python tools/prof_stage.py [n_windows] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_pkg, make_batch  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg = load_pkg()
wins, prob, pp, tp, op = make_batch(pkg, n_win, 0)
ba = pkg.SqrtBA(general_matvec=(os.environ.get('SQRTBA_GENERAL_MATVEC') == '1'),
                pipe_stages=int(os.environ.get('SQRTBA_PIPE_STAGES', '0')))
ba.set_problem_batch(prob, pp, tp, op)
ba.debug_linearize(1)
ba.debug_step(100.0)
free_obs = int((prob.pose_fixed[prob.obs_pose] == 0).sum())
for stage, name in [(0, "k_matvec"), (1, "k_linearize"), (2, "k_qr"), (3, "k_cost"), (4, "k_backsub")]:
    ms = ba.time_stage(stage, warmup=1, reps=reps)
    print(f"{name}: {ms:.4f} ms/launch", flush=True)
    if stage == 0:
        print(f"  matvec algorithmic {free_obs * 216 / 1e6:.1f} MB -> {free_obs * 216 / ms / 1e6:.1f} GB/s")
print("obs", prob.n_obs, "free_obs", free_obs, "points", prob.n_point)
ba.close()
