"""CPU tests of the N>1 host logic (gloo, world_size 2): landmark sharding covers the problem exactly once, the NCCL
id travels over torch.distributed, and per-rank results merge back in the global order.  No GPU compute here."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_landmark_shards_partition_the_problem(pkg, synth):
    prob = synth.make_problem(3, 40, 1, 1500, 7.0, loop=True, cand_halfwidth=10)
    for n in (1, 2, 3, 8):
        cuts = pkg.multi.landmark_cuts(prob, n)
        assert cuts[0] == 0 and cuts[-1] == prob.n_point and np.all(np.diff(cuts) >= 0)
        seen_obs, seen_pts, loads = 0, 0, []
        for r in range(n):
            sh, (l0, l1), (o0, o1) = pkg.multi.shard_by_landmark(prob, r, n)
            assert sh.n_pose == prob.n_pose and np.array_equal(sh.pose_fixed, prob.pose_fixed)
            np.testing.assert_array_equal(sh.point_xyz, prob.point_xyz[l0:l1])
            np.testing.assert_array_equal(sh.obs_point + l0, prob.obs_point[o0:o1])
            np.testing.assert_array_equal(sh.obs_pose, prob.obs_pose[o0:o1])
            assert sh.n_obs == 0 or (sh.obs_point.min() >= 0 and sh.obs_point.max() < sh.n_point)
            seen_obs += sh.n_obs
            seen_pts += sh.n_point
            loads.append(sh.n_obs)
        assert seen_obs == prob.n_obs and seen_pts == prob.n_point
        assert max(loads) - min(loads) <= 64          # balanced by observation count up to one long track


WORKER = textwrap.dedent("""
    import os, sys, numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
    import torch.distributed as dist
    from conftest import load_pkg
    pkg = load_pkg()
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    prob = pkg.synth.make_problem(5, 30, 1, 800, 6.0, loop=True, cand_halfwidth=8)
    shard, (l0, l1), _ = pkg.multi.shard_by_landmark(prob, rank, world)
    # the communicator id is created by rank 0 and broadcast; here any 128-byte payload exercises the plumbing
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    # each rank "solves" its shard (identity here) and the points are merged back in global order
    parts = [None] * world
    dist.all_gather_object(parts, shard.point_xyz + rank)
    merged = np.concatenate(parts, axis=0)
    cuts = pkg.multi.landmark_cuts(prob, world)
    want = prob.point_xyz.copy()
    for r in range(world):
        want[cuts[r]:cuts[r + 1]] += r
    assert np.array_equal(merged, want)
    # pose-sized partial vectors summed over ranks == the unsharded sum (what the all-reduce in the matvec relies on)
    import torch
    cnt = np.bincount(shard.obs_pose, minlength=prob.n_pose).astype(np.float64)
    t = torch.from_numpy(cnt.copy())
    dist.all_reduce(t)
    assert np.array_equal(t.numpy(), np.bincount(prob.obs_pose, minlength=prob.n_pose))
    dist.destroy_process_group()
    print('rank', rank, 'ok')
""")


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2
