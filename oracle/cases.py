"""Golden-case definitions shared by oracle/make_golden.py and tests/ -- TEST INFRASTRUCTURE.

``CASES`` maps a fixture name to a spec; ``build_inputs(spec)`` rebuilds the
seeded inputs; ``load_case(name)`` reads a committed fixture back
(inputs exactly as stored + the reference's outputs).
"""
from __future__ import annotations

import json
import os

import numpy as np

from pymn_b200.synth import make_counts, make_genesets

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_BASE = dict(S=3, n=160, G=120, K=5)

CASES = {
    # per-cell rank + normalise (utils.py:122-148)
    "normalize": dict(kind="normalize", seed=7, S=1, n=64, G=257, K=4, log1p=False, neg=True),
    # unsupervised fast path (MetaNeighborUS.py:228-387)
    "us_fast_ndn": dict(kind="us_fast", seed=11, ndn=True, **_BASE),
    "us_fast_nondn": dict(kind="us_fast", seed=12, ndn=False, **_BASE),
    "us_fast_nosym": dict(kind="us_fast", seed=13, ndn=True, sym=False, **_BASE),
    "us_fast_1v1": dict(kind="us_fast", seed=14, ndn=True, ovb=True, S=3, n=210, G=150, K=7),
    "us_fast_pval": dict(kind="us_fast", seed=15, ndn=True, p=True, **_BASE),
    "us_fast_edge": dict(kind="us_fast", seed=16, ndn=True, sym=False, S=3, n=150, G=100, K=5,
                         zero_cells=[0, 17, 301], drop_type=[1, 2], study_names=["s1", "s10", "s2"],
                         extra_genes=15, log1p=True),
    # unsupervised full-network path (utils.py:35-75, MetaNeighborUS.py:145-224)
    "us_default_ndn": dict(kind="us_default", seed=21, ndn=True, **_BASE),
    "us_default_nondn": dict(kind="us_default", seed=22, ndn=False, **_BASE),
    "us_default_nosym": dict(kind="us_default", seed=23, ndn=True, sym=False, **_BASE),
    "us_default_pval": dict(kind="us_default", seed=24, ndn=True, p=True, sym=False, **_BASE),
    "us_default_edge": dict(kind="us_default", seed=25, ndn=True, sym=False, S=3, n=120, G=90, K=4,
                            zero_cells=[3, 200], drop_type=[0, 1], study_names=["s1", "s10", "s2"],
                            extra_genes=10),
    # pretrained model (trainModel.py, MetaNeighborUS.py:391-430)
    "train_model": dict(kind="train_model", seed=31, S=2, n=150, G=110, K=6, zero_cells=[5], extra_genes=12),
    "pretrained_ndn": dict(kind="pretrained", seed=32, ndn=True, S=2, n=140, G=110, K=4, extra_genes=9,
                           train=dict(kind="train_model", seed=33, S=2, n=150, G=110, K=6, extra_genes=9)),
    "pretrained_nondn_1v1": dict(kind="pretrained", seed=34, ndn=False, ovb=True, S=2, n=140, G=110, K=6,
                                 train=dict(kind="train_model", seed=35, S=3, n=120, G=110, K=6)),
    # supervised (MetaNeighbor.py)
    "mn_default_ndn": dict(kind="supervised", seed=41, ndn=True, fast=False, S=3, n=90, G=200, K=4,
                           n_sets=6, set_lo=8, set_hi=40, all_hvg=True),
    "mn_default_nondn": dict(kind="supervised", seed=42, ndn=False, fast=False, S=3, n=90, G=200, K=4,
                             n_sets=4, set_lo=8, set_hi=40, all_hvg=True),
    "mn_fast_ndn": dict(kind="supervised", seed=43, ndn=True, fast=True, S=3, n=120, G=200, K=5,
                        n_sets=8, set_lo=8, set_hi=60, all_hvg=True, drop_type=[2, 1]),
    "mn_fast_nondn": dict(kind="supervised", seed=44, ndn=False, fast=True, S=3, n=120, G=200, K=5,
                          n_sets=5, set_lo=8, set_hi=60, all_hvg=True),
    # variableGenes (variableGenes.py) -- counts, and log-normalised floats (float64 X: the reference then
    # computes np.var in float64 like the kernel does)
    "var_genes_counts": dict(kind="var_genes", seed=51, S=3, n=301, G=400, K=5),
    "var_genes_log1p": dict(kind="var_genes", seed=52, S=2, n=250, G=320, K=4, log1p=True),
}


def build_inputs(spec: dict) -> dict:
    x, s, c = make_counts(spec["seed"], spec["S"], spec["n"], spec["G"], spec["K"])
    rng = np.random.default_rng(spec["seed"] + 1000)
    if spec.get("study_names"):
        names = spec["study_names"]
        s = np.array([names[int(v[5:]) - 1] for v in s], dtype=object)
    if spec.get("drop_type"):
        si, ti = spec["drop_type"]
        st_names = sorted(set(s.tolist()), key=lambda v: list(s).index(v))
        keep = ~((s == st_names[si]) & (c == f"ct{ti:03d}"))
        x, s, c = x[keep], s[keep], c[keep]
    if spec.get("log1p"):
        tot = np.maximum(x.sum(axis=1, keepdims=True), 1.0)
        x = np.log1p(x / tot * 1e3).astype(np.float32)
    if spec.get("neg"):
        x = (x - np.float32(1.0)).astype(np.float32)              # negatives + a big tie block at -1
        x[::3] *= np.float32(-0.5)
    for r in spec.get("zero_cells", []):
        x[r % x.shape[0]] = 0
    n_extra = spec.get("extra_genes", 0)
    g = x.shape[1]
    hvg = np.ones(g + n_extra, bool)
    if n_extra:
        extra = rng.poisson(1.0, size=(x.shape[0], n_extra)).astype(np.float32)
        pos = np.sort(rng.choice(g + n_extra, size=n_extra, replace=False))
        hvg[pos] = False
        full = np.empty((x.shape[0], g + n_extra), np.float32)
        full[:, hvg] = x
        full[:, ~hvg] = extra
        x = full
    genes = np.array([f"g{i:04d}" for i in range(x.shape[1])], dtype=object)
    perm = rng.permutation(x.shape[1])                            # gene names not in sorted order
    genes = genes[perm]
    out = dict(X=x, study=s, ctype=c, genes=genes, hvg=hvg)
    if spec["kind"] == "supervised":
        n_sets = spec["n_sets"]
        universe = np.concatenate([genes[: int(0.9 * len(genes))],
                                   np.array([f"zz{i}" for i in range(7)], dtype=object)])
        universe = universe[rng.permutation(len(universe))]
        gs = make_genesets(spec["seed"] + 2000, len(universe), n_sets, spec["set_lo"], spec["set_hi"])
        gs = np.concatenate([gs, np.zeros((len(universe), 1), np.int8)], axis=1)     # an empty set (dropped)
        out["genesets"] = gs
        out["gs_genes"] = universe
    return out


def storable_inputs(inp: dict) -> dict:
    d = dict(inp)
    if np.all(d["X"] == np.rint(d["X"])) and d["X"].min() >= 0 and d["X"].max() < 65535:
        d["X"] = d["X"].astype(np.uint16)
    return d


def load_case(name: str):
    """-> (spec, inputs, outputs) from the committed fixture."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
    spec = json.loads(str(z["spec"]))
    inp, out = {}, {}
    for k in z.files:
        if k.startswith("in_"):
            v = z[k]
            inp[k[3:]] = v.astype(object) if v.dtype.kind == "U" else v
        elif k.startswith("out_"):
            out[k[4:]] = z[k]
    inp["X"] = inp["X"].astype(np.float32)
    return spec, inp, out


def list_cases(kind=None):
    return [k for k, v in CASES.items() if kind is None or v["kind"] == kind]
