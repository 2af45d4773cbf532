"""NumPy emulation of the full-network path's rank -> weight map (pass 1 histogram + CDF LUT + pass 2
lookup of pymn_b200/csrc/gemm_tcgen05.cu) -- TEST INFRASTRUCTURE.

The kernels never see the exact global rank of the N^2 correlations (utils.py:44-51): they bin
rho into `nbins` bins, keep the 16-bit position inside the bin, and read the weight from a CDF
table.  This module restates that arithmetic on the CPU so that the *scheme* (bin count,
tie handling) can be held against the oracle's exact ranking without a GPU, and so that GPU tests
can check the kernels against the same integers.
"""
from __future__ import annotations

import numpy as np

from oracle import mn_oracle as O

W_ONE = 1 << 20
Q_FRAC = 16
DELTA_MAX = 2047


def planes(x):
    """Centred doubled ranks d = 2*rank - (G+1) (int64) and float32 1/||d|| (0 = zero-variance cell)."""
    r2 = O.rank2_rows(np.asarray(x))
    g = r2.shape[1]
    d = r2 - (g + 1)
    ss = (d.astype(np.float64) ** 2).sum(axis=1)
    inv = np.where(ss > 0, 1.0 / np.sqrt(np.maximum(ss, 1e-300)), 0.0).astype(np.float32)
    return d, inv


def bin_scale(nbins):
    return np.float32(np.float32(0.5) * np.float32(nbins) - np.float32(0.81))


def bin_offset(nbins):
    return np.float32(np.float32(0.5) * np.float32(nbins) + np.float32(0.043))


def q_matrix(d, inv, nbins, acc=None):
    """q_ij = floor((rho * bin_scale + bin_offset) * 2^16) as the kernels compute it in float32 (rho_q in
    gemm_tcgen05.cu): acc * (inv_i * scale) rounded, then fmaf with inv_j and the offset.  Full matrix;
    the kernels evaluate the upper triangle (i < j) only."""
    if acc is None:
        acc = (d @ d.T)
    acc = acc.astype(np.float32)                        # exact integers while |.| < 2^24
    inv_i_s = (inv * bin_scale(nbins)).astype(np.float32)
    t = (acc * inv_i_s[:, None]).astype(np.float32)
    pos = (t.astype(np.float64) * inv[None, :].astype(np.float64) + float(bin_offset(nbins))).astype(np.float32)   # fmaf
    q = np.floor(np.maximum(pos.astype(np.float32) * np.float32(1 << Q_FRAC), 0).astype(np.float64))
    return np.minimum(q, float((nbins << Q_FRAC) - 1)).astype(np.int64)


def build_lut(hist, n_diag):
    """corr_lut_kernel: packed words (W0 << 11) | rise, [nbins + 2]."""
    nbins = hist.shape[0]
    total = float(hist.sum())
    F = 2.0 * total + n_diag
    top = F - (n_diag - 1.0) * 0.5 - 1.0
    if not top > 0:
        top = 1.0
    scale = W_ONE / top
    cum = np.concatenate([[0], np.cumsum(hist)]).astype(np.float64)
    edge = np.minimum(W_ONE, np.rint(2.0 * cum * scale)).astype(np.int64)
    w0 = edge[:-1].copy()
    rise = edge[1:] - edge[:-1]
    heavy = rise >= DELTA_MAX
    mid = np.minimum(W_ONE, np.rint((2.0 * cum[:-1] + hist - 0.5) * scale)).astype(np.int64)
    w0 = np.where(heavy, mid, w0)
    rise = np.where(heavy, 0, rise)
    words = (w0 << 11) | rise
    return np.concatenate([words, [edge[-1] << 11, int(heavy.sum())]]).astype(np.int64)


def w_from_lut(lut_words, q):
    """w_from_q: integer weight (2^20 fixed point) of bin positions q through a packed table."""
    e = np.asarray(lut_words, np.int64)[q >> Q_FRAC]
    return (e >> 11) + (((e & DELTA_MAX) * (q & ((1 << Q_FRAC) - 1))) >> Q_FRAC)


def hist_upper(q, live, nbins):
    n = q.shape[0]
    iu = np.triu_indices(n, 1)
    ok = live[iu[0]] & live[iu[1]]
    return np.bincount((q[iu] >> Q_FRAC)[ok], minlength=nbins).astype(np.int64)


def weights_fixed(x, nbins=16384):
    """Integer weights W_fix (N x N, symmetric from the upper triangle, diagonal and dropped cells 0)."""
    d, inv = planes(x)
    q = q_matrix(d, inv, nbins)
    live = inv > 0
    lut = build_lut(hist_upper(q, live, nbins), d.shape[0])
    w = np.triu(w_from_lut(lut, q), 1)
    w = w + w.T
    w[~live, :] = 0
    w[:, ~live] = 0
    return w, lut


def weights(x, nbins=16384):
    """W (float64 N x N, diagonal 1, dropped cells 0) through the emulated LUT."""
    w, _ = weights_fixed(x, nbins)
    w = w.astype(np.float64) / W_ONE
    np.fill_diagonal(w, 1.0)
    return w


def tie_merged_reference(x):
    """create_nw_spearman (utils.py:56-75) with correlations rounded to 1e-11 before the global ranking:
    float64 rounding in np.corrcoef splits blocks of mathematically equal correlations into sub-blocks
    in an order that depends on the BLAS in use; merging them shows how far the reference's own output
    is defined on heavily tied data (tiny gene sets)."""
    r = O.rank2_rows(np.asarray(x)).astype(np.float64) * 0.5
    with np.errstate(invalid="ignore", divide="ignore"):
        nw = np.atleast_2d(np.corrcoef(r))
    nw = np.round(nw, 11)
    np.fill_diagonal(nw, 1)
    fin = np.isfinite(nw)
    rk = O.rank2_rows(nw[fin]).astype(np.float64) * 0.5 - 1
    rk /= rk.max()
    nw[...] = 0.0
    nw[fin] = rk
    np.fill_diagonal(nw, 1)
    return nw


def us_default_from_w(w, s, c, ndn=True):
    """metaNeighborUS_default's vote + AUROC stage on a given weight matrix (oracle arithmetic)."""
    s = np.asarray(s, dtype=object)
    labels = O.join_labels(s, c)
    lab_uniq, lab_codes = O.encode(labels)
    n_lab = len(lab_uniq)
    sum_in = w @ O._onehot(lab_codes, n_lab)
    if ndn:
        sum_in = sum_in / w.sum(axis=0)[:, None]
    out = np.full((n_lab, n_lab), np.nan)
    lab_study = np.empty(n_lab, dtype=object)
    lab_study[lab_codes] = s
    for study in sorted(set(s.tolist())):
        rows = np.flatnonzero(s == study)
        ranks = O._nanrank_cols(sum_in[rows])
        for ct in np.flatnonzero(lab_study == study):
            pos = lab_codes[rows] == ct
            n_p = float(pos.sum())
            nn = rows.shape[0] - n_p
            with np.errstate(invalid="ignore", divide="ignore"):
                out[:, ct] = (ranks[pos].sum(axis=0) / n_p - (n_p + 1) / 2) / nn
    return out, lab_uniq
