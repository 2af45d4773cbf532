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


def _worker_cells(rank, world, port, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pymn_b200 import engine
        from pymn_b200.utils import label_layout
        rng = np.random.default_rng(seed)
        n = 37
        study = rng.choice(np.array(["s1", "s2", "s10"], dtype=object), size=n)
        ctype = rng.choice(np.array(["a", "b", "c", "d"], dtype=object), size=n)
        lay = label_layout(study, ctype)
        sh = engine.cell_shard(lay, rank, world)
        L = len(lay.row_labels)
        # the own cells, in local order, are a subsequence of the global device row order with the right label ranges
        assert np.all((sh.rows >= sh.lo) & (sh.rows < sh.hi)) and len(sh.rows) == sh.hi - sh.lo
        lab_of_input = np.empty(n, np.int64)
        lab_of_input[lay.perm] = lay.cell_label
        for l in range(L):
            assert np.all(lab_of_input[sh.rows[sh.lab_row_off[l]:sh.lab_row_off[l + 1]]] == l)
        # per-cell arrays: gather -> global device row order
        gidx = torch.from_numpy(sh.gidx)
        got = engine._gather_cells(torch.from_numpy(sh.rows.astype(np.float32)), sh, gidx)
        assert np.array_equal(got.numpy(), lay.perm.astype(np.float32))
        # votes exchange: value encodes (column, input row)
        n_train = 5
        n_max = (n_train + world - 1) // world
        out = torch.zeros((world * n_max, sh.per), dtype=torch.float32)
        for d in range(world):
            for k, col in enumerate(range(d, n_train, world)):
                out[d * n_max + k, :len(sh.rows)] = torch.from_numpy(1000.0 * col + sh.rows.astype(np.float32))
        mine = engine.exchange_votes(out, sh, n_max, gidx)
        for k, col in enumerate(range(rank, n_train, world)):
            assert np.array_equal(mine[k].numpy(), 1000.0 * col + lay.perm.astype(np.float32)), (rank, col)
        # the same in two row chunks, each with its own (asynchronous) all-to-all, blocks padded beyond `per`
        Q = 2
        clen = (sh.per + Q - 1) // Q + 3
        send = torch.zeros((Q, world * n_max, clen), dtype=torch.float32)
        for q in range(Q):
            own = sh.rows[q * clen:(q + 1) * clen].astype(np.float32)
            for d in range(world):
                for k, col in enumerate(range(d, n_train, world)):
                    send[q, d * n_max + k, :len(own)] = torch.from_numpy(1000.0 * col + own)
            if q == 0:
                recv, works = engine.exchange_votes_begin(send, q)
            else:
                engine.exchange_votes_begin(send, q, recv, works)
        mine2 = engine.exchange_votes_end(recv, works, sh, n_max, gidx)
        assert torch.equal(mine2[:len(range(rank, n_train, world))], mine[:len(range(rank, n_train, world))])
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.ones(1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("seed", [0, 1])
def test_cell_shard_and_votes_exchange_gloo(tmp_path, seed):
    """Cell-sharded centroid path: each rank owns a block of input rows, sorted locally by label; the
    all-to-all hands every rank its train columns for all cells in global device row order."""
    world = 2
    mp.spawn(_worker_cells, args=(world, _free_port(), seed, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / f"ok_{rank}.npy").exists()
