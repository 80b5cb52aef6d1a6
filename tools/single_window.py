#!/usr/bin/env python
"""Latency of ONE local-BA window (the reference's actual call shape: Optimizer::LocalBundleAdjustment on one window).

  python tools/single_window.py [--config c0|c1|c2] [--reps 5] [--pcg-mode 0|1]

Prints one JSON line: ms per two-pass local BA (device events + host wall), LM trials, CG iterations, LM iters/s, obs/s."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c0")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--pcg-mode", type=int, default=0)
    args = ap.parse_args()
    pkg = load_pkg()
    prob = {"c0": pkg.synth.config_c0, "c1": pkg.synth.config_c1, "c2": pkg.synth.config_c2}[args.config](0)
    ba = pkg.SqrtBA(pcg_mode=args.pcg_mode)
    ba.set_problem(prob)
    ms, wall = [], []
    for i in range(args.reps + 2):
        ba.reset_state()
        t0 = time.perf_counter()
        st = ba.solve_local()
        wall.append(1e3 * (time.perf_counter() - t0))
        ms.append(st["ms_total"])
    tr = ba.trace()
    best = min(ms[2:])
    print(json.dumps({"config": args.config, "pcg_mode": args.pcg_mode, "n_obs": prob.n_obs, "n_free_pose": prob.n_free,
                      "ms_device_best": best, "ms_wall_best": min(wall[2:]), "lm_trials": len(tr),
                      "cg_iters": st["cg_iters_total"], "kernel_launches": st["kernel_launches"],
                      "lm_iters_per_s": len(tr) / (best * 1e-3), "obs_per_s": len(tr) * prob.n_obs / (best * 1e-3)}), flush=True)
    ba.close()


if __name__ == "__main__":
    main()
