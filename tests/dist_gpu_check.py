"""Multi-GPU determinism check (run under torchrun on >= 2 GPUs):
the full-network AUROC table and the supervised table must be bit-identical to the 1-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pandas as pd
    import pymn_b200
    from pymn_b200 import engine
    from pymn_b200.synth import make_counts, make_genesets
    from pymn_b200.utils import label_layout
    from tests.helpers import make_adata

    x, s, c = make_counts(11, 3, 1500, 300, 6)
    lay = label_layout(s, c)
    Xd = engine.to_device(x, torch.float32)
    args = (Xd, engine._i32(lay.perm), engine._i32(lay.cell_label), engine._i32(lay.seg_off), len(lay.row_labels),
            len(lay.studies), True, engine.DEFAULT_NBINS)
    multi = engine.us_default_device(*args)[0].cpu().numpy()
    single = engine.us_default_device(*args, shard=(0, 1, None))[0].cpu().numpy()
    assert np.array_equal(multi, single, equal_nan=True), "sharded full-network table differs from the single-GPU one"

    # public API, full network: every rank uploads only its slice of the cells, planes are all-gathered
    genes0 = np.array([f"g{i}" for i in range(300)], dtype=object)
    inp0 = dict(X=x, study=s, ctype=c, genes=genes0, hvg=np.ones(300, bool))
    kw = dict(fast_version=False, symmetric_output=False, save_uns=False)
    api_multi = pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kw)
    saved_shard = engine.dist_shard
    engine.dist_shard = lambda: (0, 1, None)
    try:
        api_single = pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kw)
    finally:
        engine.dist_shard = saved_shard
    assert np.array_equal(api_multi.values, api_single.values, equal_nan=True), "sharded-ingestion table differs"
    for ovb in (False, True):                      # ... and the centroid path through the API (sharded ingestion too)
        kwf = dict(fast_version=True, one_vs_best=ovb, symmetric_output=False, save_uns=False)
        f_multi = pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kwf)
        engine.dist_shard = lambda: (0, 1, None)
        try:
            f_single = pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kwf)
        finally:
            engine.dist_shard = saved_shard
        assert np.array_equal(f_multi.values, f_single.values, equal_nan=True), f"fast API table differs (ovb={ovb})"

    # pretrained model: trainModel (cells sharded, integer all-reduce of the centroid sums), then scoring against it
    import warnings

    def _single(fn):
        engine.dist_shard = lambda: (0, 1, None)
        try:
            return fn()
        finally:
            engine.dist_shard = saved_shard

    m_multi = pymn_b200.trainModel(make_adata(inp0), "study_id", "cell_type")
    m_single = _single(lambda: pymn_b200.trainModel(make_adata(inp0), "study_id", "cell_type"))
    assert list(m_multi.columns) == list(m_single.columns)
    assert np.array_equal(m_multi.values, m_single.values, equal_nan=True), "sharded trainModel differs"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kwp = dict(trained_model=m_multi, symmetric_output=False, save_uns=False)
        p_multi = pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kwp)
        p_single = _single(lambda: pymn_b200.MetaNeighborUS(make_adata(inp0), "study_id", "cell_type", **kwp))
    assert np.array_equal(p_multi.values, p_single.values, equal_nan=True), "sharded pretrained table differs"

    # centroid (fast) path on a device-resident / host matrix: cells sharded, votes exchanged, columns all-gathered
    for ovb in (False, True):
        sharded = engine.us_fast(x, lay, True, ovb)
        saved_shard = engine.dist_shard
        engine.dist_shard = lambda: (0, 1, None)
        try:
            alone = engine.us_fast(x, lay, True, ovb)
        finally:
            engine.dist_shard = saved_shard
        assert np.array_equal(sharded["table"], alone["table"], equal_nan=True), f"fast path differs (ovb={ovb})"
        assert np.array_equal(sharded["n_pos"], alone["n_pos"])

    genes = np.array([f"g{i}" for i in range(300)], dtype=object)
    inp = dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(300, bool))
    gs = pd.DataFrame(make_genesets(5, 300, 5, 10, 120), index=genes, columns=[f"set{j}" for j in range(5)])
    tab = pymn_b200.MetaNeighbor(make_adata(inp), "study_id", "cell_type", gs, fast_version=True, save_uns=False)
    # every rank holds the full gathered table; compare with an unsharded run on rank 0's GPU
    M = sys.modules["pymn_b200.MetaNeighbor"]          # the module (the package re-exports the function)
    saved = M._dist_info
    M._dist_info = lambda: (0, 1)
    ref = pymn_b200.MetaNeighbor(make_adata(inp), "study_id", "cell_type", gs, fast_version=True, save_uns=False)
    M._dist_info = saved
    assert np.array_equal(tab.values, ref.values, equal_nan=True), "gene-set sharded table differs"
    dist.barrier()
    engine.release_peer_buffers()
    dist.barrier()
    if rank == 0:
        print(f"dist check ok on {dist.get_world_size()} GPUs: full-network, centroid, pretrained, trainModel and "
              "supervised tables bit-identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
