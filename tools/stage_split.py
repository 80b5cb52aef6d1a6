#!/usr/bin/env python
"""Per-stage device time of one solve (CUDA events around each stage; serialises the stages, so use for shares only)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c0")
ap.add_argument("--pcg-mode", type=int, default=0)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
pkg = load_pkg()
if args.config == "c3":
    prob = pkg.synth.config_c3(0, scale=args.scale, n_kf=max(int(1500 * args.scale), 160))
else:
    prob = {"c0": pkg.synth.config_c0, "c1": pkg.synth.config_c1, "c2": pkg.synth.config_c2}[args.config](0)
ba = pkg.SqrtBA(pcg_mode=args.pcg_mode, stage_timing=True)
ba.set_problem(prob)
for _ in range(2):
    ba.reset_state()
    st = ba.solve_global(10, False) if args.config == "c3" else ba.solve_local()
st["config"] = args.config
st["pcg_mode"] = args.pcg_mode
st["us_per_cg_iter"] = 1e3 * st["ms_pcg"] / max(st["cg_iters_total"], 1)
print(json.dumps(st), flush=True)
ba.close()
