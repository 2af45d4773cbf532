"""Multi-GPU bit-identity as a `-m gpu` test: launches tests/dist_gpu_check.py under torchrun on every GPU of the box
(2, 4 or 8) and requires rc 0.  The check compares, on each rank, the sharded result with an unsharded run of the
same process: full-network table (row tiles dealt to ranks, NCCL sums of integer partials, sharded ingestion),
centroid path (train columns dealt to ranks) with and without one-vs-best, supervised table (gene sets dealt to
ranks).  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_tables_bit_identical_across_gpu_counts():
    n = _n_gpus()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {n}")
    worlds = [w for w in (2, 4, 8) if w <= n]
    for w in worlds:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={w}",
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + w), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
        res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, f"{w} GPUs: rc {res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-3000:]}"
        assert f"dist check ok on {w} GPUs" in res.stdout
