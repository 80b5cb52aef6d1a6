#!/usr/bin/env python
"""Cheap stand-in for an 8-GPU global-BA run: every launched rank takes ONE shard of an `--nshards`-way landmark split
of C3, so the per-GPU workload of the persistent PCG kernel (tiles per CTA, barriers, vector phase) is what an
`--nshards`-GPU run sees, on 1 GPU (no exchange) or 2 GPUs (exchange with one peer instead of seven).

  SQRTBA_LIB=sqrtlm-slam_b200/libsqrtba_prof.so python tools/gba_proxy.py --nshards 8            # phase cycles on stderr
  python -m torch.distributed.run --nproc-per-node 2 tools/gba_proxy.py --nshards 8

The linear systems differ from the real run's (a shard alone is a different operator), so only per-iteration TIMES are
meaningful here, not iteration counts or results."""
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
    ap.add_argument("--nshards", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--kf", type=int, default=1500)
    ap.add_argument("--max-cg", type=int, default=300)
    ap.add_argument("--pcg-mode", type=int, default=0, help="5 = 6x6 block-Jacobi instead of the chunk preconditioner (A/B)")
    args = ap.parse_args()
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_pkg()
    prob = pkg.synth.config_c3(0, scale=args.scale, n_kf=args.kf)
    shard, _, _ = pkg.multi.shard_by_landmark(prob, rank, args.nshards)
    ba = pkg.SqrtBA(device=local, pcg_max_iters=args.max_cg, stage_timing=True, pcg_mode=args.pcg_mode)
    if world > 1:
        pkg.multi.init_comm(ba, rank, world)
    ba.set_problem(shard)
    out = []
    for rep in range(2):
        ba.reset_state()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = ba.solve_global(args.iters, False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append({"s": dt, "cg_iters": st["cg_iters_total"], "ms_pcg": st["ms_pcg"],
                    "us_per_cg_iter": 1e3 * st["ms_pcg"] / max(st["cg_iters_total"], 1), "ms_linearize": st["ms_linearize"],
                    "ms_qr": st["ms_qr"], "lm_trials": st["lm_trials"], "grid": st.get("reserved2")})
    free_obs = int((shard.pose_fixed[shard.obs_pose] == 0).sum())
    if rank == 0:
        print(json.dumps({"nshards": args.nshards, "ranks": world, "shard_obs": shard.n_obs, "free_obs": free_obs,
                          "ideal_us_per_matvec_at_6550GBs": free_obs * 216 / 6550.7e3, "runs": out}), flush=True)
    ba.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
