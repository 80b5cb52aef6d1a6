#!/usr/bin/env python
"""Per-kernel HBM roofline of the stage kernels on one of the BASELINE.json shapes (CUDA events, one GPU).

  python tools/stage_roofline.py --config c2|c3|c4 [--windows 256] [--scale 1.0] [--variants]

Algorithmic bytes follow SURVEY.md §8(d) / DESIGN.md §4: matvec and back-substitution 216 B per free-pose observation,
linearise 24 B in + 264 B out per observation, landmark QR 312 B per observation + 72 B per landmark, cost 24 B."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_pkg, make_batch, measured_peaks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--variants", action="store_true", help="A/B the big-window matvec knobs")
    ap.add_argument("--qr-variants", action="store_true", help="A/B the pipelined landmark-QR variants")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    pkg = load_pkg()
    peaks, kind = measured_peaks()
    batch = None
    if args.config == "c4":
        _, prob, pp, tp, op = make_batch(pkg, args.windows, 0)
        batch = (pp, tp, op)
    elif args.config == "c2":
        prob = pkg.synth.config_c2(0, scale=args.scale)
    elif args.config == "c0":
        prob = pkg.synth.config_c0(0)
    else:
        prob = pkg.synth.config_c3(0, scale=args.scale, n_kf=max(int(1500 * args.scale), 160))
    free_obs = int((prob.pose_fixed[prob.obs_pose] == 0).sum())
    variants = [dict()]
    if args.variants:
        variants += [dict(plain_qr=True),
                     dict(no_reorder=True), dict(general_matvec=True)]
    if args.qr_variants:
        variants += [dict(qr_variant=v) for v in (2, 5, 6)]
    for kw in variants:
        ba = pkg.SqrtBA(**kw)
        if batch:
            ba.set_problem_batch(prob, *batch)
        else:
            ba.set_problem(prob)
        ba.debug_linearize(1)
        ba.debug_step(100.0)
        # SURVEY 8(d) figures (canonical layout); 5 = fused linearise + QR: 264 B per observation + 168 B per landmark
        bytes_of = {0: free_obs * 216.0, 1: prob.n_obs * 288.0, 2: prob.n_obs * 312.0 + prob.n_point * 72.0,
                    3: prob.n_obs * 24.0, 4: free_obs * 216.0 + prob.n_point * 168.0,
                    5: prob.n_obs * 264.0 + prob.n_point * 168.0}
        out = {"config": args.config, "variant": kw, "n_obs": prob.n_obs, "free_obs": free_obs, "n_point": prob.n_point,
               "n_free_pose": prob.n_free, "peak_gbs": peaks["hbm_gbs"], "peak_kind": kind, "kernels": {}}
        stages = [(0, "k_matvec"), (1, "k_linearize"), (2, "k_qr"), (3, "k_cost"), (4, "k_backsub"), (5, "k_linqr_fused")]
        for stage, name in (stages if not kw else (stages[1:3] if kw.get("plain_qr") else (stages[1:3] if "qr_variant" in kw else stages[:1]))):
            ms = ba.time_stage(stage, warmup=3, reps=args.reps)
            gbs = bytes_of[stage] / (ms * 1e-3) / 1e9
            out["kernels"][name] = {"ms": round(ms, 5), "alg_MB": round(bytes_of[stage] / 1e6, 2), "GBs": round(gbs, 1),
                                    "frac": round(gbs / peaks["hbm_gbs"], 4)}
        print(json.dumps(out), flush=True)
        ba.close()


if __name__ == "__main__":
    main()
