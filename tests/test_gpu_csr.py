"""GPU parity: CSR ingestion (scipy.sparse input) gives exactly the dense result."""
import warnings

import numpy as np
import pytest
import torch
from scipy import sparse

from oracle import mn_oracle as O
from pymn_b200.synth import make_counts
from tests.helpers import make_adata

pytestmark = pytest.mark.gpu


def _eng():
    from pymn_b200 import engine
    return engine


@pytest.mark.parametrize("n,g,kind", [(70, 300, "counts"), (33, 2500, "counts"), (40, 700, "floats"),
                                      (12, 4000, "dense_floats"), (20, 64, "big_counts")])
def test_csr_rank_equals_dense_and_oracle(n, g, kind):
    eng = _eng()
    rng = np.random.default_rng(n + g)
    x = rng.poisson(rng.gamma(0.5, 2.0, size=(1, g)), size=(n, g)).astype(np.float32)
    if kind == "floats":
        x = np.log1p(x * rng.uniform(0.5, 2.0, size=(n, 1))).astype(np.float32)     # non-integer values -> sort path
        x[::4] *= -1.0                                                                   # negatives too
    elif kind == "dense_floats":
        x = rng.normal(size=(n, g)).astype(np.float32)                                  # > 1024 non-zeros -> CTA path
        x[:, ::7] = 0
    elif kind == "big_counts":
        x = x * 1000.0                                                                   # counts >= 2048 -> sort path
    x[0] = 0
    csr = sparse.csr_matrix(x)
    csr_explicit = csr.copy().tolil()
    csr_explicit = sparse.csr_matrix(csr_explicit)
    rows = rng.permutation(n).astype(np.int32)
    a = eng.rank_normalize(eng.to_device_matrix(csr), row_index=eng._i32(rows), want_rank2=True)
    b = eng.rank_normalize(eng.to_device(x), row_index=eng._i32(rows), want_rank2=True)
    torch.cuda.synchronize()
    assert np.array_equal(a.rank2.cpu().numpy(), O.rank2_rows(x[rows]))
    assert np.array_equal(a.rank2.cpu().numpy(), b.rank2.cpu().numpy())
    assert np.array_equal(a.d_hi.cpu().numpy(), b.d_hi.cpu().numpy())
    assert np.array_equal(a.inv_norm.cpu().numpy(), b.inv_norm.cpu().numpy())
    assert np.array_equal(a.flags.cpu().numpy() & 7, b.flags.cpu().numpy() & 7)


def test_api_sparse_input_matches_dense():
    import pymn_b200
    x, s, c = make_counts(21, 3, 400, 300, 6)
    genes = np.array([f"g{i}" for i in range(300)], dtype=object)
    dense = dict(X=x, study=s, ctype=c, genes=genes, hvg=np.arange(300) % 5 != 0)
    ad_d = make_adata(dense)
    ad_s = make_adata(dense)
    ad_s.X = sparse.csr_matrix(x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for kw in (dict(fast_version=True), dict(fast_version=True, one_vs_best=True), dict(fast_version=False)):
            td = pymn_b200.MetaNeighborUS(ad_d, "study_id", "cell_type", symmetric_output=False, save_uns=False, **kw)
            ts = pymn_b200.MetaNeighborUS(ad_s, "study_id", "cell_type", symmetric_output=False, save_uns=False, **kw)
            assert list(td.index) == list(ts.index)
            assert np.array_equal(td.values, ts.values, equal_nan=True), kw
        md = pymn_b200.trainModel(ad_d, "study_id", "cell_type")
        ms = pymn_b200.trainModel(ad_s, "study_id", "cell_type")
        assert np.array_equal(md.values, ms.values)


def test_csc_column_tiles():
    """mn_csc_gather_dense: a gene set's columns of a sparse matrix as a dense tile (explicit column list and range)."""
    eng = _eng()
    rng = np.random.default_rng(3)
    x = rng.poisson(0.3, size=(257, 90)).astype(np.float32)
    csc = eng.to_device_csc(sparse.csr_matrix(x))
    cols = np.array([5, 0, 89, 17, 17, 42, 3], dtype=np.int32)
    got = eng.csc_columns_dense(csc, eng._i32(cols)).cpu().numpy()
    assert np.array_equal(got, x[:, cols])
    assert np.array_equal(eng.csc_columns_dense(csc, None, 13, 50).cpu().numpy(), x[:, 13:63])


def test_supervised_sparse_input_matches_dense():
    """MetaNeighbor on a sparse AnnData.X (gene-set sub-setting through CSC column tiles, MetaNeighbor.py:80-93)
    gives exactly the dense tables, in both modes; the gene sets address adata.var_names in their own order."""
    import pandas as pd
    import pymn_b200
    from pymn_b200.synth import make_genesets
    x, s, c = make_counts(33, 3, 300, 240, 5)
    genes = np.array([f"g{i}" for i in range(240)], dtype=object)
    inp = dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(240, bool))
    ad_d = make_adata(inp)
    ad_s = make_adata(inp)
    ad_s.X = sparse.csr_matrix(x)
    gs = pd.DataFrame(make_genesets(7, 240, 6, 8, 100), index=genes[::-1].copy(), columns=[f"set{j}" for j in range(6)])
    for fast in (True, False):
        td = pymn_b200.MetaNeighbor(ad_d, "study_id", "cell_type", gs, fast_version=fast, save_uns=False)
        ts = pymn_b200.MetaNeighbor(ad_s, "study_id", "cell_type", gs, fast_version=fast, save_uns=False)
        assert list(td.index) == list(ts.index) and list(td.columns) == list(ts.columns)
        assert np.array_equal(td.values, ts.values, equal_nan=True), f"fast={fast}"
    # a gene universe much larger than the sets: the host cuts the matrix down first (same result)
    small = gs.iloc[:60]
    t1 = pymn_b200.MetaNeighbor(ad_d, "study_id", "cell_type", small, fast_version=True, save_uns=False)
    t2 = pymn_b200.MetaNeighbor(ad_s, "study_id", "cell_type", small, fast_version=True, save_uns=False)
    assert np.array_equal(t1.values, t2.values, equal_nan=True)


def test_variable_genes_sparse_matches_dense():
    """variableGenes on a sparse AnnData.X (variableGenes.py:109-135): per-gene medians / variances from CSC column
    tiles equal the dense ones, hence the same gene selection."""
    import pymn_b200
    eng = _eng()
    x, s, c = make_counts(35, 3, 500, 700, 4)
    genes = np.array([f"g{i}" for i in range(700)], dtype=object)
    inp = dict(X=x, study=s, ctype=c, genes=genes, hvg=np.ones(700, bool))
    ad_d = make_adata(inp)
    ad_s = make_adata(inp)
    ad_s.X = sparse.csr_matrix(x)
    vd = pymn_b200.variableGenes(ad_d, "study_id", return_vect=True)
    vs = pymn_b200.variableGenes(ad_s, "study_id", return_vect=True)
    assert np.array_equal(np.asarray(vd), np.asarray(vs))
    n = x.shape[0]
    perm, seg = np.arange(n, dtype=np.int32), np.array([0, n // 2, n], dtype=np.int32)
    md, vd2 = eng.gene_stats(x, perm, seg)
    ms, vs2 = eng.gene_stats(sparse.csc_matrix(x), perm, seg)
    assert np.array_equal(md, ms) and np.array_equal(vd2, vs2)
