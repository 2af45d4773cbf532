"""Seeded synthetic scRNA-like inputs of the BASELINE.json shapes.

Per cell type k a gene-mean vector mu_k ~ Gamma(0.5, 2); per study a
multiplicative batch effect e_s ~ LogNormal(0, 0.3) per gene; each cell's mean
is mixed with 20-40 % of the neighbouring type (so clusters overlap and AUROCs
are not all exactly 1); counts ~ Poisson.  About 70-90 % of the entries are
zero.  Labels are Python ``str`` (never ``category``).

``make_counts`` is the NumPy generator used by the tests, the golden fixtures
and the CPU baseline; ``make_counts_torch`` draws the same distribution on a
CUDA device for the BASELINE-scale bench shapes (not bit-identical to the
NumPy stream -- the bench compares GPU and CPU on one and the same tensor).
"""
from __future__ import annotations

import numpy as np


def make_labels(n_studies: int, cells_per_study: int, n_types: int):
    """Study / cell-type label vectors: cells of a study are contiguous, types
    are assigned round-robin inside a study."""
    study = np.repeat(np.arange(n_studies), cells_per_study)
    ctype = np.tile(np.arange(cells_per_study) % n_types, n_studies)
    s = np.array([f"study{v + 1}" for v in study], dtype=object)
    c = np.array([f"ct{v:03d}" for v in ctype], dtype=object)
    return s, c, study, ctype


def make_counts(seed: int, n_studies: int, cells_per_study: int, n_genes: int, n_types: int,
                mix=(0.2, 0.4), dtype=np.float32):
    """Returns (X [N x G] dtype, study str[N], celltype str[N])."""
    rng = np.random.default_rng(seed)
    mu = rng.gamma(0.5, 2.0, size=(n_types, n_genes))
    eff = rng.lognormal(0.0, 0.3, size=(n_studies, n_genes))
    s, c, study, ctype = make_labels(n_studies, cells_per_study, n_types)
    n = study.shape[0]
    m = rng.uniform(mix[0], mix[1], size=n)[:, None]
    lam = ((1 - m) * mu[ctype] + m * mu[(ctype + 1) % n_types]) * eff[study]
    x = rng.poisson(lam).astype(dtype)
    return x, s, c


def make_counts_torch(seed: int, n_studies: int, cells_per_study: int, n_genes: int, n_types: int,
                      device, chunk: int = 32768, mix=(0.2, 0.4)):
    """Same distribution drawn on ``device`` in row chunks; returns
    (X float32 [N x G] on device, study codes int64[N] (cpu), type codes int64[N] (cpu))."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    mu = torch.as_tensor(rng.gamma(0.5, 2.0, size=(n_types, n_genes)), dtype=torch.float32, device=device)
    eff = torch.as_tensor(rng.lognormal(0.0, 0.3, size=(n_studies, n_genes)), dtype=torch.float32, device=device)
    _, _, study, ctype = make_labels(n_studies, cells_per_study, n_types)
    n = study.shape[0]
    x = torch.empty((n, n_genes), dtype=torch.float32, device=device)
    st = torch.as_tensor(study, device=device)
    ct = torch.as_tensor(ctype, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = torch.rand((hi - lo, 1), generator=g, device=device) * (mix[1] - mix[0]) + mix[0]
        lam = ((1 - m) * mu[ct[lo:hi]] + m * mu[(ct[lo:hi] + 1) % n_types]) * eff[st[lo:hi]]
        x[lo:hi] = torch.poisson(lam, generator=g)
    return x, study, ctype


def make_genesets(seed: int, n_genes: int, n_sets: int, lo: int = 10, hi: int = 200):
    """Random 0/1 gene-set membership matrix (genes x sets), set sizes ~ U{lo..hi}."""
    rng = np.random.default_rng(seed)
    gs = np.zeros((n_genes, n_sets), dtype=np.int8)
    for j in range(n_sets):
        k = int(rng.integers(lo, min(hi, n_genes) + 1))
        gs[rng.choice(n_genes, size=k, replace=False), j] = 1
    return gs
