"""Drop-in ``MetaNeighborUS`` (reference pymn/MetaNeighborUS.py:12-142).

Same signature, defaults, assertion messages, warnings and ``adata.uns``
outputs as the reference; the arithmetic runs on the GPU (``engine``).
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from . import engine
from .utils import check_gene_count, label_layout, prepare_matrix, select_genes, with_gc_paused


def _pvals(roc: np.ndarray, n_pos: np.ndarray, n_neg: np.ndarray) -> np.ndarray:
    """Normal approximation of Mann-Whitney U, one-sided (utils.py:180-188);
    element-wise post-processing of the L x L table on the host."""
    from scipy import stats
    n_pos = np.asarray(n_pos, np.float64).reshape(-1, 1)
    n_neg = np.asarray(n_neg, np.float64).reshape(-1, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        U = roc * n_pos * n_neg
        Z = np.abs(U - n_pos * n_neg / 2) / np.sqrt(n_pos * n_neg * (n_pos + n_neg + 1) / 12)
    return stats.norm.sf(Z)


def _keep_rows(n_pos):
    return np.flatnonzero(np.asarray(n_pos) > 0)


def metaNeighborUS_fast(X, S, C, node_degree_normalization, one_vs_best, compute_p):
    """MetaNeighborUS.py:228-306.  S, C: pandas Series (or arrays) of labels."""
    S = S.values if hasattr(S, "values") else np.asarray(S)
    C = C.values if hasattr(C, "values") else np.asarray(C)
    Xd = engine.start_upload(prepare_matrix(X))         # the upload runs while the host encodes the labels
    lay = label_layout(S, C)
    r = engine.us_fast(Xd, lay, bool(node_degree_normalization), bool(one_vs_best))
    keep = _keep_rows(r["n_pos"])        # labels whose cells were all zero-variance vanish (:273-281)
    labels = lay.row_labels[keep]
    tab = pd.DataFrame(r["table"][np.ix_(keep, keep)], index=labels, columns=labels)
    if compute_p:
        n_pos = r["n_pos"][keep].astype(np.float64)
        seg_tot = np.array([r["n_pos"][lay.seg_lab_off[s]:lay.seg_lab_off[s + 1]].sum() for s in lay.label_study[keep]],
                           dtype=np.float64)
        p = _pvals(r["auroc"][np.ix_(keep, keep)], n_pos, seg_tot - n_pos)
        return tab, pd.DataFrame(p, index=labels, columns=labels)
    return tab


def MetaNeighborUS_from_trained(trained_model, test_data, study_col, ct_col, node_degree_normalization,
                                one_vs_best, compute_p):
    """MetaNeighborUS.py:391-430."""
    # (the obs columns go to label_layout as they are: an Arrow-backed pandas array is dictionary-encoded without a
    #  detour through a million Python strings -- 150 ms at config 5)
    S = study_col.values if hasattr(study_col, "values") else study_col
    C = ct_col.values if hasattr(ct_col, "values") else ct_col
    Xd = engine.start_upload(prepare_matrix(test_data))         # the upload runs while the host encodes the labels
    lay = label_layout(S, C, replace_bar=True)
    cols = trained_model.columns
    cent = np.ascontiguousarray(trained_model.iloc[1:].values.T, dtype=np.float32)      # [L_train, G]
    n_cells = np.asarray(trained_model.iloc[0].values, dtype=np.float32)
    st_names = np.array([str(c).split("|")[0] for c in cols.values], dtype=object)    # :341-343
    st_uniq, st_codes = np.unique(st_names, return_inverse=True)
    r = engine.us_pretrained(Xd, lay, cent, n_cells, st_codes.astype(np.int32), len(st_uniq),
                             bool(node_degree_normalization), bool(one_vs_best))
    keep = _keep_rows(r["n_pos"])
    labels = lay.row_labels[keep]
    tab = pd.DataFrame(r["table"][keep], index=labels, columns=cols)
    if compute_p:
        n_pos = r["n_pos"][keep].astype(np.float64)
        seg_tot = np.array([r["n_pos"][lay.seg_lab_off[s]:lay.seg_lab_off[s + 1]].sum() for s in lay.label_study[keep]],
                           dtype=np.float64)
        p = _pvals(r["auroc"][keep], n_pos, seg_tot - n_pos)
        return tab, pd.DataFrame(p, index=labels, columns=cols)
    return tab


def metaNeighborUS_default(adata, study_col, ct_col, node_degree_normalization, compute_p):
    """MetaNeighborUS.py:145-224: full cell x cell network.  Output is indexed
    [train label, test label] in globally sorted label order, like the reference."""
    S = adata.obs[study_col].values         # numpy or pandas array; label_layout factorises either as it is
    C = adata.obs[ct_col].values
    # the upload (and, on several GPUs, K1 + the plane exchange) runs while the host encodes the labels
    Xd = engine.start_upload(prepare_matrix(adata.X), full_network=True)
    lay = label_layout(S, C)
    r = engine.us_default(Xd, lay, bool(node_degree_normalization))
    order = np.argsort(lay.sorted_pos)                       # row-order ids in globally sorted order
    labels = lay.row_labels[order]
    auc = r["auroc"][np.ix_(order, order)]                   # [test, train]
    tab = pd.DataFrame(auc.T, index=labels, columns=labels)  # -> [train, test]  (:213)
    if compute_p:
        n_pos = r["n_pos"][order].astype(np.float64)
        seg_tot = np.array([lay.seg_off[s + 1] - lay.seg_off[s] for s in lay.label_study[order]], dtype=np.float64)
        p = _pvals(auc, n_pos, seg_tot - n_pos)
        return tab, pd.DataFrame(p.T, index=labels, columns=labels)
    return tab


@with_gc_paused
def MetaNeighborUS(adata,
                   study_col,
                   ct_col,
                   var_genes="highly_variable",
                   symmetric_output=True,
                   node_degree_normalization=True,
                   fast_version=False,
                   one_vs_best=False,
                   trained_model=None,
                   save_uns=True,
                   compute_p=False,
                   mn_key="MetaNeighborUS"):
    """Unsupervised MetaNeighbor; see reference pymn/MetaNeighborUS.py:12-60 for
    the argument documentation (identical here)."""
    assert study_col in adata.obs_keys(), "Study Col not in adata"
    assert ct_col in adata.obs_keys(), "Cluster Col not in adata"

    if trained_model is not None:
        var_genes = adata.var_names[np.isin(adata.var_names, trained_model.index)]
        trained_model = pd.concat([pd.DataFrame(trained_model.iloc[0]).T, trained_model.loc[var_genes]])
    elif type(var_genes) is str:
        assert (var_genes in adata.var_keys()
                ), f"If passing a string ({var_genes}) for var names, it must be in adata.var_keys()"
        var_genes = adata.var_names[np.asarray(adata.var[var_genes].values, dtype=bool)]
    else:
        var_genes = adata.var_names[np.isin(adata.var_names, var_genes)]

    assert var_genes.shape[0] > 2, "Must have at least 2 genes"
    check_gene_count(var_genes.shape[0])
    if var_genes.shape[0] < 5:
        warnings.warn("You should have at least 5 Variable Genes", category=UserWarning)
    if one_vs_best:
        assert (fast_version or trained_model is not None), "If you want to run in one_vs_best mode you must also \
         run in fast version mode or use a pretrained model"

    if trained_model is None:
        # (hash-based unique: np.unique would sort 10^5-10^6 label strings just to count them)
        assert (pd.unique(adata.obs[study_col].values).shape[0] > 1), "Need more than 1 study"
        if fast_version:
            assert (adata.obs[study_col].dtype.name != "category"
                    ), "Study Col is a category type, cast to either string or int"
            assert (adata.obs[ct_col].dtype.name != "category"
                    ), "Cell Type Col is a category type, cast to either string or int"
            cell_nv = metaNeighborUS_fast(select_genes(adata, var_genes).X, adata.obs[study_col], adata.obs[ct_col],
                                          node_degree_normalization, one_vs_best, compute_p)
        else:
            cell_nv = metaNeighborUS_default(select_genes(adata, var_genes), study_col, ct_col, node_degree_normalization,
                                             compute_p)
    else:
        cell_nv = MetaNeighborUS_from_trained(trained_model, select_genes(adata, var_genes).X, adata.obs[study_col].values,
                                              adata.obs[ct_col].values, node_degree_normalization, one_vs_best,
                                              compute_p)
    if compute_p:
        cell_p = cell_nv[1]
        cell_nv = cell_nv[0]
        cell_p = cell_p.astype(float)

    cell_nv = cell_nv.astype(float)
    if symmetric_output and not one_vs_best:
        cell_nv = (cell_nv + cell_nv.T) / 2
    if save_uns:
        if one_vs_best:
            adata.uns[f"{mn_key}_1v1"] = cell_nv
        else:
            adata.uns[mn_key] = cell_nv
        adata.uns[f"{mn_key}_params"] = {
            "fast": fast_version,
            "node_degree_normalization": node_degree_normalization,
            "study_col": study_col,
            "ct_col": ct_col,
            "one_vs_best": one_vs_best,
            "symmetric_output": symmetric_output,
        }
        if compute_p:
            adata.uns[f"{mn_key}_pval"] = cell_p
    else:
        return cell_nv
