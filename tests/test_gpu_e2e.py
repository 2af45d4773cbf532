"""GPU parity tests, end to end through the public API (the call a pyMN user makes).

1. every committed golden fixture (outputs of the unmodified reference);
2. larger seeded cases against the oracle (sizes the oracle finishes in seconds),
   including BASELINE config 1 at full size;
3. size-independent properties at bigger shapes.

Tolerance: AUROC tables within max(1e-5, 2/(n_pos*n_neg)) of the
float64 reference -- one swap of two adjacent votes moves a cell by exactly
1/(n_pos*n_neg); NaN patterns must match exactly.
"""
import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import cases as C
from oracle import mn_oracle as O
from pymn_b200.synth import make_counts, make_genesets
from tests.helpers import assert_table_close, make_adata, make_genesets_df

pytestmark = pytest.mark.gpu


def _tol(inp, flips=2.5):
    """max(1e-5, flips / (n_pos * n_neg)) with the smallest label / study sizes of the case."""
    s, c = np.asarray(inp["study"], dtype=object), np.asarray(inp["ctype"], dtype=object)
    lab = O.join_labels(s, c)
    worst = 0.0
    for st in set(s.tolist()):
        n_t = int((s == st).sum())
        for l in set(lab[s == st].tolist()):
            n_p = int((lab == l).sum())
            if 0 < n_p < n_t:
                worst = max(worst, flips / (n_p * (n_t - n_p)))
    return max(1e-5, worst)


def _run_us(inp, spec):
    import pymn_b200
    ad = make_adata(inp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=(spec["kind"] == "us_fast"),
                                 node_degree_normalization=spec["ndn"], one_vs_best=spec.get("ovb", False),
                                 symmetric_output=spec.get("sym", True), compute_p=spec.get("p", False))
    return ad


@pytest.mark.parametrize("name", C.list_cases("us_fast") + C.list_cases("us_default"))
def test_golden_unsupervised(name):
    spec, inp, ref = C.load_case(name)
    ad = _run_us(inp, spec)
    key = "MetaNeighborUS_1v1" if spec.get("ovb") else "MetaNeighborUS"
    tab = ad.uns[key]
    flips = 2.5 if spec["kind"] == "us_fast" else 4.5
    assert_table_close(tab.values, tab.index, tab.columns, ref["auroc_values"], ref["auroc_rows"], ref["auroc_cols"],
                       _tol(inp, flips), name, skip_cols=spec.get("noise_cols"))
    import json
    assert ad.uns["MetaNeighborUS_params"] == json.loads(str(ref["params"]))
    if spec.get("p"):
        p = ad.uns["MetaNeighborUS_pval"]
        got, want = np.asarray(p.values, float), ref["pval_values"]
        assert [str(v) for v in p.index] == [str(v) for v in ref["pval_rows"]]
        # p-values are a steep function of the AUROC: compare on the z-scale loosely
        assert np.allclose(np.log10(np.maximum(got, 1e-300)), np.log10(np.maximum(want, 1e-300)), atol=0.35)


def test_golden_train_model():
    import pymn_b200
    spec, inp, ref = C.load_case("train_model")
    model = pymn_b200.trainModel(make_adata(inp), "study_id", "cell_type")
    assert_table_close(model.values, model.index, model.columns, ref["model_values"], ref["model_rows"],
                       ref["model_cols"], 2e-5, "train_model")
    assert np.array_equal(model.values[0], ref["model_values"][0])          # n_cells exact


@pytest.mark.parametrize("name", C.list_cases("pretrained"))
def test_golden_pretrained(name):
    import pymn_b200
    spec, inp, ref = C.load_case(name)
    model = pd.DataFrame(ref["model_values"], index=[str(v) for v in ref["model_rows"]],
                         columns=[str(v) for v in ref["model_cols"]])
    ad = make_adata(inp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tab = pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", trained_model=model,
                                       node_degree_normalization=spec["ndn"], one_vs_best=spec.get("ovb", False),
                                       symmetric_output=False, save_uns=False)
    assert_table_close(tab.values, tab.index, tab.columns, ref["auroc_values"], ref["auroc_rows"], ref["auroc_cols"],
                       _tol(inp), name)


@pytest.mark.parametrize("name", C.list_cases("supervised"))
def test_golden_supervised(name):
    import pymn_b200
    spec, inp, ref = C.load_case(name)
    tab = pymn_b200.MetaNeighbor(make_adata(inp), "study_id", "cell_type", make_genesets_df(inp),
                                 node_degree_normalization=spec["ndn"], fast_version=spec["fast"], save_uns=False)
    # mean over 3 studies of cells each within `flips` quanta
    assert_table_close(tab.values, tab.index, tab.columns, ref["auroc_values"], ref["auroc_rows"], ref["auroc_cols"],
                       _tol(inp, 2.5 if spec["fast"] else 4.5), name)


# ---- larger cases against the oracle ----------------------------------------
def _inputs(seed, S, n, G, K):
    x, s, c = make_counts(seed, S, n, G, K)
    genes = np.array([f"g{i}" for i in range(G)], dtype=object)
    return dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(G, bool))


@pytest.mark.parametrize("ovb", [False, True])
def test_config1_full_size_fast(ovb):
    """BASELINE config 1: 2 studies x 10k cells x 2,000 HVGs x 20 cell types, fast_version=True."""
    inp = _inputs(1, 2, 10000, 2000, 20)
    spec = dict(kind="us_fast", ndn=True, ovb=ovb, sym=False)
    ad = _run_us(inp, spec)
    tab = ad.uns["MetaNeighborUS_1v1" if ovb else "MetaNeighborUS"]
    ref = O.metaneighbor_us_fast(inp["X"], inp["study"], inp["ctype"], True, ovb)
    d = assert_table_close(tab.values, tab.index, tab.columns, ref[0], ref[1], ref[2], 1e-5, "config1")
    print(f"config1 ovb={ovb}: max |dAUROC| = {d:.3e}")


def test_fast_g3000_hi_lo_planes():
    """G > 2048 exercises the two-plane (hi + lo) operand path of the tensor-core GEMM."""
    inp = _inputs(4, 3, 1500, 3000, 12)
    ad = _run_us(inp, dict(kind="us_fast", ndn=True, sym=False))
    tab = ad.uns["MetaNeighborUS"]
    ref = O.metaneighbor_us_fast(inp["X"], inp["study"], inp["ctype"], True, False)
    assert_table_close(tab.values, tab.index, tab.columns, ref[0], ref[1], ref[2], _tol(inp), "g3000")


@pytest.mark.parametrize("ndn", [True, False])
def test_default_path_medium(ndn):
    """Full-network path, N = 3,000 (9e6 network entries ranked globally on both sides)."""
    inp = _inputs(2, 3, 1000, 600, 5)
    ad = _run_us(inp, dict(kind="us_default", ndn=ndn, sym=False))
    tab = ad.uns["MetaNeighborUS"]
    ref = O.metaneighbor_us_default(inp["X"], inp["study"], inp["ctype"], ndn)
    d = assert_table_close(tab.values, tab.index, tab.columns, ref[0], ref[1], ref[2], _tol(inp, 2.5), "default")
    print(f"default ndn={ndn}: max |dAUROC| = {d:.3e} (quantum {1.0 / (200 * 800):.2e})")


def test_supervised_medium():
    import pymn_b200
    x, s, c = make_counts(3, 3, 500, 400, 5)
    genes = np.array([f"g{i}" for i in range(400)], dtype=object)
    inp = dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(400, bool))
    gs = make_genesets(9, 400, 4, 10, 200)
    gdf = pd.DataFrame(gs, index=genes, columns=[f"set{j}" for j in range(4)])
    for fast in (False, True):
        tab = pymn_b200.MetaNeighbor(make_adata(inp), "study_id", "cell_type", gdf, fast_version=fast, save_uns=False)
        ref = O.metaneighbor_supervised(x, genes, s, c, gs, genes, np.array(gdf.columns, dtype=object), True, fast)
        assert_table_close(tab.values, tab.index, tab.columns, ref[0], ref[1], ref[2], _tol(inp, 2.5), f"sup fast={fast}")


# ---- size-independent properties at larger shapes -----------------------------
def test_default_path_properties_20k():
    """N = 20,000: (a) the histogram holds every pair exactly once, (b) two runs are
    bit-identical (integer accumulation), (c) matching cell types replicate."""
    from pymn_b200 import engine
    from pymn_b200.utils import label_layout
    x, s, c = make_counts(7, 2, 10000, 512, 8)
    lay = label_layout(s, c)
    r1 = engine.us_default(x, lay, True)
    r2 = engine.us_default(x, lay, True)
    n = x.shape[0]
    assert r1["hist_total"] == n * (n - 1) // 2
    assert np.array_equal(r1["auroc"], r2["auroc"], equal_nan=True)
    L = len(lay.row_labels)
    same_type = np.array([[a.split("|")[1] == b.split("|")[1] for b in lay.row_labels] for a in lay.row_labels])
    other_study = lay.label_study[:, None] != lay.label_study[None, :]
    assert r1["auroc"][same_type & other_study].min() > 0.9
    assert np.all((r1["auroc"] >= 0) & (r1["auroc"] <= 1))


@pytest.mark.parametrize("shuffle", [False, True])
def test_default_path_streamed_upload_identical(shuffle, monkeypatch):
    """The streamed pipeline (row chunks uploaded last-first on a side stream, K1 + pass 1 per chunk while
    earlier rows are still in flight) gives the same table, bit for bit, as upload-then-compute --
    for study-sorted cells (the case that overlaps) and for an arbitrary cell order."""
    import torch
    from pymn_b200 import engine
    from pymn_b200.utils import label_layout
    x, s, c = make_counts(11, 3, 4200, 64, 5)
    if shuffle:
        order = np.random.default_rng(0).permutation(x.shape[0])
        x, s, c = x[order], s[order], c[order]
    xh = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).pin_memory()
    lay = label_layout(s, c)
    ref = engine.us_default(xh.numpy().copy(), lay, True)
    monkeypatch.setattr(engine, "STREAM_MIN_BYTES", 0)
    monkeypatch.setattr(engine, "STREAM_CHUNKS", 4)
    up = engine.start_upload(xh.numpy())
    assert isinstance(up, engine.AsyncUpload) and len(up.chunks) == 4
    got = engine.us_default(up, lay, True)
    assert np.array_equal(got["auroc"], ref["auroc"]) and np.array_equal(got["n_pos"], ref["n_pos"])
    assert got["hist_total"] == ref["hist_total"] and np.array_equal(got["flags"], ref["flags"])


@pytest.mark.parametrize("shuffle", [False, True])
def test_default_path_staged_pageable_upload_identical(shuffle, monkeypatch):
    """A PAGEABLE matrix goes through the worker-thread staging (StagedUpload: pieces of the pinned staging buffers,
    last rows first) and the same streamed pipeline; the table is identical to upload-then-compute.  The fast path
    (wait for the whole matrix) is checked on the same upload object type."""
    from pymn_b200 import engine
    from pymn_b200.utils import label_layout
    x, s, c = make_counts(12, 3, 4200, 64, 5)
    if shuffle:
        order = np.random.default_rng(1).permutation(x.shape[0])
        x, s, c = x[order], s[order], c[order]
    x = np.ascontiguousarray(x, dtype=np.float32)
    lay = label_layout(s, c)
    ref = engine.us_default(x.copy(), lay, True)
    ref_fast = engine.us_fast(x.copy(), lay, True, True)
    monkeypatch.setattr(engine, "STREAM_MIN_BYTES", 0)
    monkeypatch.setattr(engine, "STREAM_CHUNKS", 5)
    monkeypatch.setattr(engine, "STAGE_BYTES", 100 * 64 * 4)          # 100 rows per staged piece
    engine._STAGE.clear()
    try:
        up = engine.start_upload(x)
        assert isinstance(up, engine.StagedUpload)
        got = engine.us_default(up, lay, True)
        assert np.array_equal(got["auroc"], ref["auroc"]) and np.array_equal(got["n_pos"], ref["n_pos"])
        assert got["hist_total"] == ref["hist_total"] and np.array_equal(got["flags"], ref["flags"])
        got_fast = engine.us_fast(engine.start_upload(x), lay, True, True)
        assert np.array_equal(got_fast["table"], ref_fast["table"], equal_nan=True)
    finally:
        engine._STAGE.clear()


# ---- variableGenes -----------------------------------------------------------
@pytest.mark.parametrize("name", C.list_cases("var_genes"))
def test_golden_variable_genes(name):
    """The boolean gene vector equals the unmodified reference's, gene for gene."""
    import pymn_b200
    spec, inp, ref = C.load_case(name)
    ad = make_adata(inp)
    got = pymn_b200.variableGenes(ad, "study_id", return_vect=True)
    assert [str(g) for g in got.index] == [str(g) for g in ref["genes"]]
    assert np.array_equal(np.asarray(got.values, bool), ref["selected"])
    assert pymn_b200.variableGenes(ad, "study_id") is None
    assert np.array_equal(np.asarray(ad.var["highly_variable"].values, bool), ref["selected"])


def test_gene_stats_kernel_exact():
    """mn_gene_stats: medians are exact order statistics (odd and even study sizes, negatives, ties,
    a gene count that is not a multiple of 32), variances match float64 NumPy to rounding."""
    from pymn_b200 import engine
    rng = np.random.default_rng(3)
    n, g = 4097, 77
    x = rng.poisson(0.7, size=(n, g)).astype(np.float32)
    x[:, 5] = rng.normal(size=n).astype(np.float32)            # floats incl. negatives
    x[:, 6] = -x[:, 6]
    x[:, 7] = 3.0                                              # constant gene
    seg = np.array([0, 1000, 1001, 1001, 4097], dtype=np.int32)     # even, single-cell, EMPTY, odd studies
    perm = rng.permutation(n).astype(np.int32)
    med, var = engine.gene_stats(x, perm, seg)
    xs = x[perm].astype(np.float64)
    for k in range(4):
        blk = xs[seg[k]:seg[k + 1]]
        if blk.shape[0] == 0:
            assert np.isnan(med[k]).all() and np.isnan(var[k]).all()
            continue
        assert np.array_equal(med[k], np.median(blk, axis=0)), f"median, study {k}"
        assert np.allclose(var[k], np.var(blk, axis=0), rtol=1e-12, atol=1e-12), f"variance, study {k}"


def test_variable_genes_config_size_matches_oracle():
    """20,000 cells x 2,000 genes, 2 studies: same genes as the NumPy restatement."""
    import pymn_b200
    x, s, c = make_counts(5, 2, 10000, 2000, 20)
    genes = np.array([f"g{i}" for i in range(x.shape[1])], dtype=object)
    ad = make_adata(dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(x.shape[1], bool)))
    got = pymn_b200.variableGenes(ad, "study_id", return_vect=True)
    ref = O.variable_genes(x, s)
    assert np.array_equal(np.asarray(got.values, bool), ref)


def test_pretrained_model_through_csv(tmp_path):
    """trainModel -> save_model -> load_model -> MetaNeighborUS(trained_model=...) gives the
    table of the in-memory model, bit for bit (the protocol-2 round trip)."""
    import pymn_b200
    spec, inp, ref = C.load_case("pretrained_ndn")
    train_inp = C.build_inputs(spec["train"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = pymn_b200.trainModel(make_adata(train_inp), "study_id", "cell_type")
        pymn_b200.save_model(model, tmp_path / "m.csv")
        loaded = pymn_b200.load_model(tmp_path / "m.csv")
        assert np.array_equal(loaded.values, model.values) and list(loaded.columns) == list(model.columns)
        kw = dict(node_degree_normalization=True, symmetric_output=False, save_uns=False)
        a = pymn_b200.MetaNeighborUS(make_adata(inp), "study_id", "cell_type", trained_model=model, **kw)
        b = pymn_b200.MetaNeighborUS(make_adata(inp), "study_id", "cell_type", trained_model=loaded, **kw)
    assert np.array_equal(a.values, b.values, equal_nan=True) and list(a.index) == list(b.index)
    assert_table_close(b.values, b.index, b.columns, ref["auroc_values"], ref["auroc_rows"], ref["auroc_cols"],
                       _tol(inp), "pretrained via csv")
