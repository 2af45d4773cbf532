"""Minimal stand-in for ``anndata.AnnData``.

``anndata`` is not installable in the build image, and the hot path only
touches a handful of AnnData attributes: ``X``, ``obs``,
``var``, ``uns``, ``var_names``, ``obs_names``, ``obs_keys()``,
``var_keys()``, ``uns_keys()`` and ``adata[rows, cols]`` slicing with bool
masks, name arrays or a pandas Index (MetaNeighborUS.py:66-77,101,107,113;
MetaNeighbor.py:66,81-83; trainModel.py:30-37).  The public entry points of
this package accept a real ``anndata.AnnData`` or this class alike -- they
only use the attributes listed above.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy import sparse


class AnnDataLite:
    def __init__(self, X, obs: pd.DataFrame, var: pd.DataFrame | None = None, uns: dict | None = None):
        n, g = X.shape
        if var is None:
            var = pd.DataFrame(index=pd.Index([f"g{i}" for i in range(g)], dtype=object))
        assert obs.shape[0] == n and var.shape[0] == g, "obs/var do not match X"
        self.X = X
        self.obs = obs
        self.var = var
        self.uns = {} if uns is None else uns

    # -- attribute surface the hot path reads --------------------------------
    @property
    def shape(self):
        return self.X.shape

    @property
    def var_names(self) -> pd.Index:
        return self.var.index

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    def obs_keys(self):
        return list(self.obs.columns)

    def var_keys(self):
        return list(self.var.columns)

    def uns_keys(self):
        return list(self.uns.keys())

    # -- slicing --------------------------------------------------------------
    @staticmethod
    def _resolve(sel, names: pd.Index):
        if isinstance(sel, slice):
            return np.arange(len(names))[sel]
        if isinstance(sel, (pd.Series, pd.Index)):
            sel = sel.values
        sel = np.asarray(sel)
        if sel.dtype == bool:
            return np.flatnonzero(sel)
        if np.issubdtype(sel.dtype, np.integer):
            return sel
        idx = names.get_indexer(sel)
        if (idx < 0).any():
            raise KeyError("names not found in index")
        return idx

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key, slice(None))
        rows = self._resolve(key[0], self.obs.index)
        cols = self._resolve(key[1], self.var.index)
        X = self.X
        if sparse.issparse(X):
            X = X.tocsr()[rows][:, cols]
        else:
            X = np.asarray(X)[np.ix_(rows, cols)]
        return AnnDataLite(X, self.obs.iloc[rows], self.var.iloc[cols], self.uns)
