"""Worker of tests/test_gpu_multi.py (one rank per GPU under torchrun): landmark-sharded global BA on N GPUs checked
against the CPU ORACLE (not against the 1-GPU run): trial sequence, per-trial cost, lambda, pose RMS, points.

  python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_worker.py [--scale S] [--kf K] [--robust R] [--iters I]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.3)
    ap.add_argument("--kf", type=int, default=450)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--robust", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from conftest import load_pkg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_pkg()
    prob = pkg.synth.config_c3(args.seed, scale=args.scale, n_kf=args.kf)  # same seed on every rank: identical problem
    shard, (l0, l1), _ = pkg.multi.shard_by_landmark(prob, rank, world)
    ba = pkg.SqrtBA(device=local)
    pkg.multi.init_comm(ba, rank, world)
    ba.set_problem(shard)
    st = ba.solve_global(args.iters, bool(args.robust))
    tr, poses, pts = ba.trace(), ba.poses(), ba.points()
    parts = [None] * world
    dist.all_gather_object(parts, (tr, poses, pts))
    ok = True
    out = {"n_gpus": world, "n_pose": prob.n_pose, "n_obs": prob.n_obs, "persistent_pcg": st["persistent_pcg"],
           "peer_exchange": st["peer_exchange"], "chunk_precond": st["chunk_precond"], "lm_trials": len(tr)}
    if rank == 0:
        from oracle import refba
        from test_gpu_parity import COST_RTOL, POSE_R_RMS, POSE_T_RMS, pose_rms
        ref = refba.RefBA(prob, threads=max(1, min(os.cpu_count() or 1, 32)))
        ref.solve_global(args.iters, bool(args.robust))
        rt = ref.trace()
        try:
            # every rank holds the same poses and the same LM trace (replicated CG state, rank-ordered sums)
            for r in range(1, world):
                assert np.array_equal(parts[r][0][:, :8], tr[:, :8]), f"rank {r}: LM trace differs from rank 0"
                assert np.array_equal(parts[r][1], poses), f"rank {r}: poses differ from rank 0"
            assert len(tr) == len(rt), (len(tr), len(rt))
            assert np.array_equal(tr[:, [0, 1, 2, 7]], rt[:, [0, 1, 2, 7]])
            np.testing.assert_allclose(tr[:, 4], rt[:, 4], rtol=COST_RTOL)
            np.testing.assert_allclose(tr[:, 5], rt[:, 5], rtol=COST_RTOL)
            np.testing.assert_allclose(tr[:, 3], rt[:, 3], rtol=1e-5)
            t_rms, r_rms = pose_rms(poses, ref.poses(), prob.pose_fixed == 0)
            out["pose_t_rms_m"], out["pose_r_rms_rad"] = t_rms, r_rms
            assert t_rms <= POSE_T_RMS and r_rms <= POSE_R_RMS, (t_rms, r_rms)
            pts_all = np.concatenate([p[2] for p in parts], axis=0)
            np.testing.assert_allclose(pts_all, ref.points(), rtol=1e-5, atol=1e-4)
            out["final_chi2"], out["oracle_final_chi2"] = float(tr[-1, 5]), float(rt[-1, 5])
            if world > 1:
                assert st["persistent_pcg"] == 1 and st["peer_exchange"] == 1, st
        except AssertionError as exc:
            ok = False
            out["error"] = repr(exc)[:2000]
        out["ok"] = ok
        print("MULTI_GPU_PARITY " + json.dumps(out), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ba.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
