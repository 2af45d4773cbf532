"""Shared helpers for the test-suite: run the oracle / the product on a golden
case through the same top-level steps the reference's API wrappers perform
(HVG sub-setting, model gene alignment, symmetrisation), and compare tables."""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mn_oracle as O          # noqa: E402  (tests are allowed to import the oracle)
from pymn_b200.anndata_lite import AnnDataLite   # noqa: E402


def make_adata(inp) -> AnnDataLite:
    obs = pd.DataFrame({"study_id": np.asarray(inp["study"], dtype=object),
                        "cell_type": np.asarray(inp["ctype"], dtype=object)},
                       index=pd.Index([f"c{i}" for i in range(inp["X"].shape[0])], dtype=object))
    var = pd.DataFrame({"highly_variable": np.asarray(inp["hvg"], bool)},
                       index=pd.Index(np.asarray(inp["genes"], dtype=object), dtype=object))
    return AnnDataLite(np.asarray(inp["X"], np.float32), obs, var)


def make_genesets_df(inp) -> pd.DataFrame:
    gs = inp["genesets"]
    return pd.DataFrame(gs, index=pd.Index(np.asarray(inp["gs_genes"], dtype=object), dtype=object),
                        columns=[f"set{j}" for j in range(gs.shape[1])])


def oracle_run(spec, inp, train_inp=None):
    """Outputs of the oracle for a case, keyed like the golden fixtures."""
    kind = spec["kind"]
    hv = np.asarray(inp["hvg"], bool)
    x = np.asarray(inp["X"], np.float32)[:, hv]
    s, c = inp["study"], inp["ctype"]
    out = {}
    if kind == "normalize":
        out["xnorm"] = O.normalize_cells(x)
        out["rank2"] = O.rank2_rows(x)
    elif kind == "us_fast":
        r = O.metaneighbor_us_fast(x, s, c, spec["ndn"], spec.get("ovb", False), spec.get("p", False))
        tab = r[0]
        if spec.get("sym", True) and not spec.get("ovb", False):
            tab = (tab + tab.T) / 2
        out.update(auroc_values=tab, auroc_rows=r[1], auroc_cols=r[2])
        if spec.get("p", False):
            out.update(pval_values=r[3], pval_rows=r[1], pval_cols=r[2])
    elif kind == "us_default":
        r = O.metaneighbor_us_default(x, s, c, spec["ndn"], spec.get("p", False))
        tab = r[0]
        if spec.get("sym", True):
            tab = (tab + tab.T) / 2
        out.update(auroc_values=tab, auroc_rows=r[1], auroc_cols=r[2])
        if spec.get("p", False):
            out.update(pval_values=r[3], pval_rows=r[1], pval_cols=r[2])
    elif kind == "train_model":
        model, cols = O.train_model(x, s, c)
        out.update(model_values=model, model_cols=cols,
                   model_rows=np.array(["n_cells"] + list(np.asarray(inp["genes"], dtype=object)[hv]), dtype=object))
    elif kind == "pretrained":
        thv = np.asarray(train_inp["hvg"], bool)
        model, mcols = O.train_model(np.asarray(train_inp["X"], np.float32)[:, thv],
                                     train_inp["study"], train_inp["ctype"])
        mgenes = np.asarray(train_inp["genes"], dtype=object)[thv]
        # MetaNeighborUS.py:65-70: query genes (in query order) that are in the model index
        qgenes = np.asarray(inp["genes"], dtype=object)
        pos = {g: i for i, g in enumerate(mgenes.tolist())}
        sel = np.array([g in pos for g in qgenes.tolist()])
        aligned = np.vstack([model[:1], model[1:][[pos[g] for g in qgenes[sel].tolist()]]])
        r = O.metaneighbor_us_from_trained(aligned, mcols, np.asarray(inp["X"], np.float32)[:, sel], s, c,
                                           spec["ndn"], spec.get("ovb", False))
        out.update(auroc_values=r[0], auroc_rows=r[1], auroc_cols=r[2])
    elif kind == "var_genes":
        out["selected"] = O.variable_genes(np.asarray(inp["X"], np.float32), s)
        out["genes"] = np.asarray(inp["genes"], dtype=object)
    elif kind == "supervised":
        gs = inp["genesets"]
        names = np.array([f"set{j}" for j in range(gs.shape[1])], dtype=object)
        tab, labels, kept = O.metaneighbor_supervised(np.asarray(inp["X"], np.float32), inp["genes"], s, c, gs,
                                                      inp["gs_genes"], names, spec["ndn"], spec["fast"])
        out.update(auroc_values=tab, auroc_rows=labels, auroc_cols=kept)
    else:
        raise ValueError(kind)
    return out


def assert_table_close(got_vals, got_rows, got_cols, ref_vals, ref_rows, ref_cols, atol, what="", skip_cols=None):
    """``skip_cols``: column labels left out of the comparison (a spec's ``noise_cols``: columns whose
    reference values are decided by float64 rounding noise, documented where the case is defined)."""
    assert [str(v) for v in got_rows] == [str(v) for v in ref_rows], f"{what}: row labels differ"
    assert [str(v) for v in got_cols] == [str(v) for v in ref_cols], f"{what}: column labels differ"
    got_vals = np.array(got_vals, np.float64)
    ref_vals = np.array(ref_vals, np.float64)
    assert got_vals.shape == ref_vals.shape, f"{what}: shape {got_vals.shape} vs {ref_vals.shape}"
    if skip_cols:
        sel = np.array([str(v) in set(skip_cols) for v in ref_cols])
        got_vals[:, sel] = np.nan
        ref_vals[:, sel] = np.nan
    assert np.array_equal(np.isnan(got_vals), np.isnan(ref_vals)), f"{what}: NaN pattern differs"
    d = np.nanmax(np.abs(got_vals - ref_vals)) if np.isfinite(ref_vals).any() else 0.0
    assert d <= atol, f"{what}: max |diff| {d:.3e} > {atol:.1e}"
    return d
