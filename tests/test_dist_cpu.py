"""world_size-2 gloo tests (CPU) of the multi-process host logic:
gene-set column sharding + all-gather, and the integer all-reduce hook of the
full-network path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_sets, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pymn_b200 import engine
        from pymn_b200.MetaNeighbor import _dist_info, gather_columns, shard_columns
        assert _dist_info() == (rank, world)
        K = 4
        mine = shard_columns(n_sets, rank, world)
        local = np.stack([np.arange(K) * 100.0 + j for j in mine], axis=1) if len(mine) else np.zeros((K, 0))
        local[0, :] = np.nan                                  # NaN entries must survive the gather
        full = gather_columns(local, n_sets, rank, world)
        r, w, reduce_fn = engine.dist_shard()
        assert (r, w) == (rank, world) and reduce_fn is not None
        t = torch.full((5,), rank + 1, dtype=torch.int64)
        reduce_fn(t)
        np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
        np.save(os.path.join(out_dir, f"sum_{rank}.npy"), t.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_sets", [7, 2, 1])
def test_geneset_sharding_gloo(tmp_path, n_sets):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_sets, str(tmp_path)), nprocs=world, join=True)
    K = 4
    want = np.stack([np.arange(K) * 100.0 + j for j in range(n_sets)], axis=1)
    want[0, :] = np.nan
    for rank in range(world):
        got = np.load(tmp_path / f"full_{rank}.npy")
        assert np.array_equal(got, want, equal_nan=True)
        assert np.array_equal(np.load(tmp_path / f"sum_{rank}.npy"), np.full(5, 3))


def _worker_rows(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pymn_b200 import engine
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)       # every rank holds the same host matrix
        lo, hi, per = engine.shard_rows(n, rank, world)
        got = engine.gather_rows(full[lo:hi].clone(), per, world)           # what us_default_sharded does with the planes
        vec = engine.gather_rows(full[lo:hi, 0].clone(), per, world)
        np.save(os.path.join(out_dir, f"rows_{rank}.npy"), got.numpy())
        np.save(os.path.join(out_dir, f"vec_{rank}.npy"), vec.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 2])
def test_sharded_ingestion_row_gather_gloo(tmp_path, n):
    """Data-parallel ingestion: each rank contributes its row slice; after the all-gather row j of the buffer is
    input row j on every rank (ragged last block padded), so `index_select(perm)` yields the label-sorted order."""
    world = 2
    mp.spawn(_worker_rows, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    want = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    for rank in range(world):
        got = np.load(tmp_path / f"rows_{rank}.npy")
        assert got.shape[0] >= n and np.array_equal(got[:n], want) and not got[n:].any()
        assert np.array_equal(np.load(tmp_path / f"vec_{rank}.npy")[:n], want[:, 0])
