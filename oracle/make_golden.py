"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- build container only.

TEST INFRASTRUCTURE.  Usage (from the repo root, in the build container):

    python -m oracle.make_golden            # writes tests/golden/<case>.npz

Each fixture holds the inputs (counts as uint16, labels as str arrays, the
kwargs as JSON) and the reference's outputs (table values + row / column
labels).  The reference runs through ``oracle/ref_loader.py`` (its own .py
files, SciPy-shimmed ``bottleneck``).  Case definitions live in
``oracle/cases.py`` so that the tests rebuild the very same inputs.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases as C                      # noqa: E402
from oracle.ref_loader import load_reference       # noqa: E402
from pymn_b200.anndata_lite import AnnDataLite     # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _adata(inp):
    obs = pd.DataFrame({"study_id": inp["study"].astype(object), "cell_type": inp["ctype"].astype(object)},
                       index=pd.Index([f"c{i}" for i in range(inp["X"].shape[0])], dtype=object))
    var = pd.DataFrame({"highly_variable": inp["hvg"].astype(bool)},
                       index=pd.Index(inp["genes"].astype(object), dtype=object))
    return AnnDataLite(inp["X"].astype(np.float32), obs, var)


def _table(df):
    return dict(values=np.asarray(df.values, np.float64),
                rows=np.asarray([str(v) for v in df.index], dtype=str),
                cols=np.asarray([str(v) for v in df.columns], dtype=str))


def run_reference(name, spec):
    ref = load_reference()
    inp = C.build_inputs(spec)
    out = {}
    kind = spec["kind"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "normalize":
            xn = ref.utils.normalize_cells(inp["X"].astype(np.float32))
            import scipy.stats
            out["xnorm"] = np.asarray(xn, np.float64)
            out["rank2"] = np.rint(2 * scipy.stats.rankdata(inp["X"], "average", axis=1)).astype(np.int32)
        elif kind in ("us_fast", "us_default"):
            ad = _adata(inp)
            res = ref.MetaNeighborUS.MetaNeighborUS(
                ad, "study_id", "cell_type", fast_version=(kind == "us_fast"),
                node_degree_normalization=spec["ndn"], one_vs_best=spec.get("ovb", False),
                symmetric_output=spec.get("sym", True), compute_p=spec.get("p", False), save_uns=True)
            assert res is None
            key = "MetaNeighborUS_1v1" if spec.get("ovb", False) else "MetaNeighborUS"
            out.update({f"auroc_{k}": v for k, v in _table(ad.uns[key]).items()})
            if spec.get("p", False):
                out.update({f"pval_{k}": v for k, v in _table(ad.uns["MetaNeighborUS_pval"]).items()})
            out["params"] = np.asarray(json.dumps(ad.uns["MetaNeighborUS_params"], sort_keys=True))
        elif kind == "train_model":
            ad = _adata(inp)
            model = ref.trainModel.trainModel(ad, "study_id", "cell_type")
            out.update({f"model_{k}": v for k, v in _table(model).items()})
        elif kind == "pretrained":
            ad_ref = _adata(C.build_inputs(spec["train"]))
            model = ref.trainModel.trainModel(ad_ref, "study_id", "cell_type")
            ad = _adata(inp)
            res = ref.MetaNeighborUS.MetaNeighborUS(
                ad, "study_id", "cell_type", trained_model=model, node_degree_normalization=spec["ndn"],
                one_vs_best=spec.get("ovb", False), symmetric_output=False, save_uns=False)
            out.update({f"auroc_{k}": v for k, v in _table(res).items()})
            out.update({f"model_{k}": v for k, v in _table(model).items()})
        elif kind == "var_genes":
            ad = _adata(inp)
            ad.X = inp["X"].astype(np.float64)
            res = ref.variableGenes.variableGenes(ad, "study_id", return_vect=True)
            out["selected"] = np.asarray(res.values, bool)
            out["genes"] = np.asarray([str(v) for v in res.index], dtype=str)
        elif kind == "supervised":
            ad = _adata(inp)
            gs = pd.DataFrame(inp["genesets"], index=pd.Index(inp["gs_genes"].astype(object), dtype=object),
                              columns=[f"set{j}" for j in range(inp["genesets"].shape[1])])
            res = ref.MetaNeighbor.MetaNeighbor(ad, "study_id", "cell_type", gs,
                                                node_degree_normalization=spec["ndn"],
                                                fast_version=spec["fast"], save_uns=False)
            out.update({f"auroc_{k}": v for k, v in _table(res).items()})
        else:
            raise ValueError(kind)
    return inp, out


def main_big(only):
    """BASELINE-shaped cases: outputs + input digests only (oracle/cases.py BIG_CASES)."""
    import time
    for name, spec in C.BIG_CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        inp, out = run_reference(name, spec)
        payload = {f"out_{k}": v for k, v in out.items() if not k.startswith("model_")}
        payload["spec"] = np.asarray(json.dumps(spec, sort_keys=True))
        payload["in_sha256"] = np.asarray(C.inputs_digest(inp))
        if "train" in spec:
            payload["train_sha256"] = np.asarray(C.inputs_digest(C.build_inputs(spec["train"])))
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: wrote {path} ({os.path.getsize(path)} bytes, reference took {time.time() - t0:.0f} s)", flush=True)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "--big" in sys.argv[1:]:
        return main_big(set(a for a in sys.argv[1:] if a != "--big"))
    only = set(sys.argv[1:])
    for name, spec in C.CASES.items():
        if only and name not in only:
            continue
        inp, out = run_reference(name, spec)
        payload = {f"in_{k}": (v.astype(str) if v.dtype == object else v) for k, v in C.storable_inputs(inp).items()}
        payload.update({f"out_{k}": v for k, v in out.items()})
        payload["spec"] = np.asarray(json.dumps(spec, sort_keys=True))
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
