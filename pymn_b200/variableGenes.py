"""Drop-in ``variableGenes`` (reference pymn/variableGenes.py:7-77): genes whose variance is in the top
quartile of their median-expression decile in EVERY study.

The per-study, per-gene medians and variances -- the only part that touches the N x G matrix -- come
from one GPU kernel (``mn_gene_stats``: 4-pass radix select + float64 moments); the decile / quartile
binning of the two G-vectors is the reference's own NumPy sequence on the host.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import engine
from .utils import _factorize_sorted, _stable_argsort_small, to_dense_f32


def _matrix(X):
    """A sparse AnnData.X stays sparse (CSC on the device, variableGenes.py:109-135); anything else dense float32."""
    from scipy import sparse
    return X if sparse.issparse(X) else to_dense_f32(X)


def _select(median: np.ndarray, variance: np.ndarray) -> np.ndarray:
    """variableGenes.py:28-37 on one study's gene statistics."""
    bins = np.quantile(median, q=np.linspace(0, 1, 11), method="midpoint")
    digits = np.digitize(median, bins, right=True)
    selected = np.zeros_like(digits)
    for i in np.unique(digits):
        filt = digits == i
        var_tmp = variance[filt]
        bins_tmp = np.nanquantile(var_tmp, q=np.linspace(0, 1, 5))
        g = np.digitize(var_tmp, bins_tmp)
        selected[filt] = (g >= 4).astype(float)
    return selected.astype(bool)


def compute_var_genes(adata, return_vect=True):
    """Variable genes of a single dataset (variableGenes.py:7-42)."""
    X = _matrix(adata.X)
    n = X.shape[0]
    med, var = engine.gene_stats(X, np.arange(n, dtype=np.int32), np.array([0, n], dtype=np.int32))
    sel = _select(med[0], var[0])
    if return_vect:
        return sel
    adata.var["highly_variable"] = sel


def variableGenes(adata, study_col, return_vect=False):
    """Variable genes across datasets (variableGenes.py:45-77): the intersection over studies.
    All studies are handled by one kernel launch over the study-sorted cell order."""
    assert study_col in adata.obs_keys(), "Study Col not in obs data"
    studies, codes = _factorize_sorted(adata.obs[study_col].values)
    perm = _stable_argsort_small(codes, len(studies)).astype(np.int32)
    seg_off = np.concatenate([[0], np.cumsum(np.bincount(codes, minlength=len(studies)))]).astype(np.int32)
    med, var = engine.gene_stats(_matrix(adata.X), perm, seg_off)
    genes = adata.var_names
    var_genes_mat = pd.DataFrame(index=genes)
    for k, study in enumerate(studies):
        var_genes_mat.loc[:, study] = _select(med[k], var[k])
    var_genes = np.all(var_genes_mat, axis=1)
    if return_vect:
        return var_genes
    adata.var["highly_variable"] = var_genes
