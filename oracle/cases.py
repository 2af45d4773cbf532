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
    # one-vs-best when a study holds a SINGLE cell type: its one-vs-rest AUROCs are 0/0 = NaN, so
    # compute_1v1_aurocs skips every column of that study (utils.py:207-208) and find_top_candidates' candidates[1]
    # (utils.py:231) is never reached -- the rows of that study stay NaN.  noise_cols: as a TRAIN column that
    # cluster is its study's only one, so votes / node degree = (x.C + n) / (x.C + n) = 1 for every cell up to
    # float64 rounding; the reference's one-vs-best cells of that column are decided by that rounding noise.
    "us_fast_1v1_single": dict(kind="us_fast", seed=17, ndn=True, ovb=True, S=3, n=150, G=24, K=5, only_type=[2, 1],
                               noise_cols=["study3|ct001"]),
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
    # a cell that is CONSTANT AND NON-ZERO over a gene set survives score_low_mem's sum > 0 filter
    # (MetaNeighbor.py:121), normalises to a NaN row and poisons every vote of that gene set (Appendix C-6):
    # the whole column of the result is NaN.  Cell 7 is 2.0 everywhere except on three genes.
    "mn_fast_poison": dict(kind="supervised", seed=45, ndn=True, fast=True, S=3, n=100, G=160, K=4,
                           n_sets=10, set_lo=6, set_hi=50, all_hvg=True, poison_cell=7, poison_keep_genes=[3, 50, 120]),
    # variableGenes (variableGenes.py) -- counts, and log-normalised floats (float64 X: the reference then
    # computes np.var in float64 like the kernel does)
    "var_genes_counts": dict(kind="var_genes", seed=51, S=3, n=301, G=400, K=5),
    "var_genes_log1p": dict(kind="var_genes", seed=52, S=2, n=250, G=320, K=4, log1p=True),
}


# Cases at the shapes of the BASELINE.json configs, too big to commit their inputs:
# the fixture holds the reference's OUTPUT table plus a SHA-256 of the seeded inputs, which
# ``load_big_case`` rebuilds with ``build_inputs`` and verifies.  Generated like the small ones by
# ``python -m oracle.make_golden --big`` (the unmodified reference, build container only).
BIG_CASES = {
    # config 2's regime -- full network, G = 2,000, L = 40 -- at the sizes the reference can still hold in RAM
    # (N = 4,000 is the sample bench.py's CPU arm times; N = 16,000 needs ~10 GB and ~3 min of CPU)
    "big_default_c2_4k": dict(kind="us_default", seed=2, ndn=True, sym=False, S=4, n=1000, G=2000, K=10),
    "big_default_c2_16k": dict(kind="us_default", seed=2, ndn=True, sym=False, S=4, n=4000, G=2000, K=10),
    # config 4's shape: 7 studies, G = 3,000 (two fp16 planes), 700 clusters, one-vs-rest and one-vs-best
    "big_fast_c4": dict(kind="us_fast", seed=4, ndn=True, sym=False, S=7, n=5000, G=3000, K=100),
    "big_fast_c4_1v1": dict(kind="us_fast", seed=4, ndn=True, ovb=True, sym=False, S=7, n=5000, G=3000, K=100),
    # config 5's shape: a 150-column model (3 studies x 50 clusters), 20,000 query cells with 50 labels
    "big_pretrained_c5": dict(kind="pretrained", seed=5, ndn=True, S=1, n=20000, G=3000, K=50,
                              train=dict(kind="train_model", seed=6, S=3, n=2500, G=3000, K=50)),
}


def inputs_digest(inp: dict) -> str:
    """SHA-256 over the arrays a case is built from (X bytes, labels, gene names, HVG mask)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(inp["X"], dtype=np.float32).tobytes())
    for k in ("study", "ctype", "genes"):
        h.update("\x1f".join(str(v) for v in inp[k]).encode())
    h.update(np.asarray(inp["hvg"], bool).tobytes())
    return h.hexdigest()


def load_big_case(name: str):
    """-> (spec, inputs rebuilt from the seed and verified against the stored digest, reference outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
    spec = json.loads(str(z["spec"]))
    inp = build_inputs(spec)
    assert inputs_digest(inp) == str(z["in_sha256"]), f"{name}: the seeded inputs differ from those the reference saw"
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    if "train" in spec:
        tr = build_inputs(spec["train"])
        assert inputs_digest(tr) == str(z["train_sha256"]), f"{name}: the seeded training inputs differ"
    return spec, inp, out


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
    if spec.get("only_type"):
        si, ti = spec["only_type"]
        st_names = sorted(set(s.tolist()), key=lambda v: list(s).index(v))
        keep = (s != st_names[si]) | (c == f"ct{ti:03d}")
        x, s, c = x[keep], s[keep], c[keep]
    if spec.get("log1p"):
        tot = np.maximum(x.sum(axis=1, keepdims=True), 1.0)
        x = np.log1p(x / tot * 1e3).astype(np.float32)
    if spec.get("neg"):
        x = (x - np.float32(1.0)).astype(np.float32)              # negatives + a big tie block at -1
        x[::3] *= np.float32(-0.5)
    for r in spec.get("zero_cells", []):
        x[r % x.shape[0]] = 0
    if "poison_cell" in spec:
        r = spec["poison_cell"]
        keep_g = spec["poison_keep_genes"]
        vals = np.array([0.0, 5.0, 7.0], np.float32)[: len(keep_g)]
        x[r] = 2.0
        x[r, keep_g] = vals
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
