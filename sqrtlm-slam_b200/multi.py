"""Host-side helpers for the landmark-sharded multi-GPU mode (one process per GPU, torch.distributed for the plumbing).

Global BA shards by LANDMARK: every rank holds all poses (replicated) and a contiguous range of map
points with all their observations; the only data-path exchange is the all-reduce inside libsqrtba (NCCL)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi, synth


def landmark_cuts(prob, nranks: int) -> np.ndarray:
    """nranks+1 landmark indices: contiguous ranges balanced by observation count."""
    ptr = prob.lm_ptr()
    target = np.linspace(0, prob.n_obs, nranks + 1)
    cuts = np.searchsorted(ptr, target, side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, prob.n_point
    return np.maximum.accumulate(cuts)


def shard_by_landmark(prob, rank: int, nranks: int):
    """The sub-problem of one rank: all poses, landmarks [l0,l1) re-indexed from 0, their observations."""
    cuts = landmark_cuts(prob, nranks)
    l0, l1 = int(cuts[rank]), int(cuts[rank + 1])
    ptr = prob.lm_ptr()
    o0, o1 = int(ptr[l0]), int(ptr[l1])
    shard = synth.Problem(prob.pose_qt, prob.pose_fixed, prob.cam, prob.point_xyz[l0:l1].copy(),
                          prob.obs_pose[o0:o1].copy(), (prob.obs_point[o0:o1] - l0).astype(np.int32),
                          prob.obs_meas[o0:o1].copy(), name=f"{prob.name}-shard{rank}of{nranks}")
    return shard, (l0, l1), (o0, o1)


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = capi.lib().sqrtba_comm_unique_id(buf)
    if rc != 0:
        raise capi.SqrtBAError(f"sqrtba_comm_unique_id failed ({rc})")
    return bytes(buf)


def init_comm(ba, rank: int, nranks: int, group=None):
    """Create the library's NCCL communicator; the 128-byte id travels over torch.distributed (any backend)."""
    import torch.distributed as dist
    obj = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    ba.comm_init(nranks, rank, obj[0])


def merge_points(shards_points, cuts):
    """Stack the per-rank point blocks back into the global order."""
    return np.concatenate(list(shards_points), axis=0)
