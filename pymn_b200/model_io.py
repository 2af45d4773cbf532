"""Pretrained-model wire format.

The reference keeps a trained model as the ``(1 + G) x L`` DataFrame ``trainModel`` returns -- row
``n_cells`` first, then one row per gene, one column per ``study|cell type`` (trainModel.py:42-48) -- and
moves it between sessions as CSV: ``model.to_csv(path)`` / ``pd.read_csv(path, index_col=0)``
(protocol 2, cells 11 and 14).  Models written by R ``MetaNeighbor::trainModel`` use the same layout
(MetaNeighborUS.py:55-56).  These two helpers are that contract, plus the checks that make a bad file
fail here instead of deep inside the scoring call.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def save_model(model: pd.DataFrame, path) -> None:
    """Write a trained model exactly as the protocol does (``DataFrame.to_csv``); float64 values are
    written with repr precision, so ``load_model`` returns them bit for bit."""
    validate_model(model)
    model.to_csv(path)


def load_model(path) -> pd.DataFrame:
    """``pd.read_csv(path, index_col=0)`` + validation.  Accepts files written by this package, by the
    reference and by R ``MetaNeighbor::trainModel`` (``write.csv``)."""
    # (round_trip: pandas' default fast float parser can be one ulp off; the protocol's plain read_csv
    #  therefore perturbs a model in its last bit, this does not)
    model = pd.read_csv(path, index_col=0, float_precision="round_trip")
    model.index = pd.Index([str(v) for v in model.index], dtype=object)
    model.columns = pd.Index([str(v) for v in model.columns], dtype=object)
    validate_model(model)
    return model.astype(np.float64)


def validate_model(model: pd.DataFrame) -> None:
    assert isinstance(model, pd.DataFrame) and model.shape[0] >= 2 and model.shape[1] >= 1, \
        "a trained model is a (1 + genes) x clusters table"
    assert str(model.index[0]) == "n_cells", "first row of a trained model must be 'n_cells' (trainModel.py:46)"
    assert all("|" in str(c) for c in model.columns), "model columns must be 'study|cell type' labels"
    assert not model.index[1:].has_duplicates, "duplicate gene names in the model"
    n = np.asarray(model.iloc[0].values, dtype=np.float64)
    assert np.all(np.isfinite(n)) and np.all(n >= 0), "n_cells must be non-negative counts"
