"""Drop-in ``trainModel`` (reference pymn/trainModel.py:4-48)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import engine
from .utils import check_gene_count, label_layout, prepare_matrix, select_genes, with_gc_paused


@with_gc_paused
def trainModel(adata, study_col, ct_col, var_genes="highly_variable"):
    """Pretrains the centroid model consumed by ``MetaNeighborUS(trained_model=...)``.

    Returns a DataFrame: row ``n_cells`` then one row per variable gene, one
    column per ``study|celltype`` (sorted), exactly the reference's layout
    (trainModel.py:42-48), so models round-trip through ``to_csv``/``read_csv``
    with either implementation.
    """
    assert study_col in adata.obs_keys(), "Study Col not in adata"
    assert ct_col in adata.obs_keys(), "Cluster Col not in adata"

    if not (isinstance(var_genes, str) and var_genes == "highly_variable"):
        var_genes = adata.var_names[np.isin(adata.var_names, var_genes)]
    else:
        var_genes = adata.var_names[np.asarray(adata.var[var_genes].values, dtype=bool)]
    assert var_genes.shape[0] > 2, "Must have at least 2 genes"
    check_gene_count(var_genes.shape[0])

    sub = select_genes(adata, var_genes)
    X = engine.start_upload(prepare_matrix(sub.X))       # the upload runs while the host encodes the labels
    lay = label_layout(sub.obs[study_col].values, sub.obs[ct_col].values)      # (Arrow-backed columns stay as they are)
    r = engine.train_model(X, lay)
    flags = r["flags"]                                   # device row order
    kept = (flags & 2) == 0                              # cells with sum > 0 (trainModel.py:36-37)
    n_cells = r["n_cells"].astype(np.float64)
    # labels are those of the kept cells only; globally sorted (pd.get_dummies)
    present = np.flatnonzero(np.bincount(lay.cell_label[kept], minlength=len(lay.row_labels)) > 0)
    present = present[np.argsort(lay.sorted_pos[present])]
    cent = r["centroids"].astype(np.float64)
    # a kept constant (zero-variance, non-zero) cell is a NaN row in the reference and
    # poisons its cluster's column; reproduce that
    poisoned = np.unique(lay.cell_label[kept & ((flags & 1) != 0)])
    n_cells[poisoned] += np.bincount(lay.cell_label[kept & ((flags & 1) != 0)], minlength=len(n_cells))[poisoned]
    cent[poisoned] = np.nan
    labels = lay.row_labels[present]
    result = pd.DataFrame(cent[present].T, index=var_genes, columns=labels)
    n = pd.DataFrame(n_cells[present][None, :], index=["n_cells"], columns=labels)
    return pd.concat([n, result])
