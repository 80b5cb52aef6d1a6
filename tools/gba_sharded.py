#!/usr/bin/env python
"""Landmark-sharded global BA on N GPUs (torchrun, one rank per GPU): time-to-converge + parity with the 1-GPU solve.

  python -m torch.distributed.run --nproc-per-node N tools/gba_sharded.py [--scale 1.0] [--kf 1500] [--iters 10] [--robust 0]

Every rank holds all poses and a contiguous landmark shard; libsqrtba all-reduces the pose-sized PCG vector per CG
iteration over NCCL.  Rank 0 also solves the whole problem alone (no communicator) and prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--kf", type=int, default=1500)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--robust", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--skip-single", action="store_true")
    ap.add_argument("--pcg-mode", type=int, default=0, help="5 = 6x6 block-Jacobi instead of the chunk preconditioner (A/B)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_pkg()
    prob = pkg.synth.config_c3(0, scale=args.scale, n_kf=args.kf)   # same seed on every rank => identical problem
    shard, (l0, l1), _ = pkg.multi.shard_by_landmark(prob, rank, world)
    ba = pkg.SqrtBA(device=local, pcg_mode=args.pcg_mode)
    if world > 1:
        pkg.multi.init_comm(ba, rank, world)
    ba.set_problem(shard)
    times = []
    for _ in range(args.reps + 1):
        ba.reset_state()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = ba.solve_global(args.iters, bool(args.robust))
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    tsec = torch.tensor([min(times[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
    tr = ba.trace()
    poses = ba.poses()
    pts = ba.points()
    out = {"workload": f"C3 global BA: {prob.n_pose} keyframes, {prob.n_point} points, {prob.n_obs} observations, "
                       f"{args.iters} iterations, robust={args.robust}",
           "n_gpus": world, "time_to_converge_s": float(tsec.item()), "lm_trials": len(tr),
           "cg_iters_total": int(tr[:, 8].sum()), "final_chi2": float(tr[-1, 5]), "matvec_launches": st["cg_iters_total"],
           "persistent_pcg": st["persistent_pcg"], "peer_exchange": st["peer_exchange"], "kernel_launches": st["kernel_launches"]}
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, pts)
        pts_all = np.concatenate(allp, axis=0)
    else:
        pts_all = pts
    if rank == 0 and not args.skip_single and world > 1:
        single = pkg.SqrtBA(device=local, pcg_mode=args.pcg_mode)
        single.set_problem(prob)
        single.solve_global(args.iters, bool(args.robust))
        ts = single.trace()
        out["single_gpu_final_chi2"] = float(ts[-1, 5])
        out["trace_rel_diff_max"] = float(np.max(np.abs(ts[:, 5] - tr[:, 5]) / ts[:, 5])) if len(ts) == len(tr) else None
        out["same_trial_sequence"] = bool(len(ts) == len(tr) and np.array_equal(ts[:, [0, 1, 2, 7]], tr[:, [0, 1, 2, 7]]))
        out["pose_t_max_diff_m"] = float(np.abs(single.poses()[:, :3] - poses[:, :3]).max())
        out["point_max_diff_m"] = float(np.abs(single.points() - pts_all).max())
        single.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    ba.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
