"""Landmark-sharded global BA on several GPUs against the ORACLE: spawns one rank per GPU
with torchrun when the box has at least two GPUs; the worker (tests/multi_gpu_worker.py) shards the landmarks, runs
sqrtba_solve_global through the C ABI with the in-kernel NVLink exchange and rank 0 compares with refba."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(nproc, extra):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("MULTI_GPU_PARITY ")]
    assert res.returncode == 0 and lines, res.stdout[-3000:] + res.stderr[-3000:]
    out = json.loads(lines[-1][len("MULTI_GPU_PARITY "):])
    assert out["ok"], out
    return out


@pytest.mark.parametrize("robust", [0, 1])
def test_sharded_global_ba_matches_oracle(robust):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    nproc = 2 if n < 4 else 4
    out = _run(nproc, ["--scale", "0.3", "--kf", "450", "--robust", str(robust), "--iters", "10"])
    assert out["n_gpus"] == nproc and out["persistent_pcg"] == 1 and out["peer_exchange"] == 1
    assert out["chunk_precond"] == 1  # the sharded PCG runs with the 20-pose chunk preconditioner (all-reduced blocks)
