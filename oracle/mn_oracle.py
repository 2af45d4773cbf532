"""CPU oracle for the MetaNeighbor scoring hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy (float64) restatement of the arithmetic of
gillislab/pyMN's hot path.  It exists so that the CUDA kernels can be checked
against it; it is NOT part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  Nothing under ``pymn_b200/`` imports it, and
the product path raises if the CUDA library is missing instead of falling
back to this code.

Parity pin: the reference ships no tests or golden vectors,
so the pin is "outputs of the reference itself run here": ``oracle/make_golden.py``
executes the reference's own ``pymn/{utils,MetaNeighborUS,MetaNeighbor,
trainModel}.py`` (unmodified, loaded by path with a SciPy shim for the absent
``bottleneck`` C extension -- see ``oracle/ref_loader.py``) on seeded inputs
and commits inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
holds this restatement to those vectors.

Third-party arithmetic restated here (not vendored in /root/reference):
``bottleneck`` (unpinned in requirements.txt:7) ``rankdata`` / ``nanrankdata``
= average-tie, 1-based ranks, float64; identical to
``scipy.stats.rankdata(method="average")``.

Every function cites the reference file:line it follows (paths relative to
/root/reference/pymn/).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "rank2_rows", "rankdata_avg", "normalize_cells", "join_labels", "encode",
    "compute_aurocs", "compute_pvals", "one_vs_best", "metaneighbor_us_fast",
    "create_nw_spearman", "metaneighbor_us_default", "train_model",
    "metaneighbor_us_from_trained", "score_low_mem", "score_default",
    "metaneighbor_supervised",
]


# --------------------------------------------------------------------------
# ranking  (bottleneck.rankdata: utils.py:45,69,141,175)
# --------------------------------------------------------------------------
def rank2_rows(x: np.ndarray) -> np.ndarray:
    """Doubled average-tie ranks along the last axis, as int64 (2*rank, exact).

    Restates ``bottleneck.rankdata(res, axis=1)`` (utils.py:141).  For a run
    of equal values occupying sorted positions first..last (0-based) the
    1-based average rank is (first+last)/2 + 1, so 2*rank = first+last+2.
    """
    x = np.asarray(x)
    if x.ndim == 1:
        return rank2_rows(x[None, :])[0]
    n, g = x.shape
    if g == 0:
        return np.zeros((n, 0), np.int64)
    order = np.argsort(x, axis=1, kind="stable")
    xs = np.take_along_axis(x, order, axis=1)
    pos = np.broadcast_to(np.arange(g, dtype=np.int64), (n, g))
    new = np.ones((n, g), bool)
    new[:, 1:] = xs[:, 1:] != xs[:, :-1]
    first = np.maximum.accumulate(np.where(new, pos, 0), axis=1)
    last_flag = np.ones((n, g), bool)
    last_flag[:, :-1] = new[:, 1:]
    last = np.minimum.accumulate(np.where(last_flag, pos, g)[:, ::-1], axis=1)[:, ::-1]
    r2s = first + last + 2
    out = np.empty((n, g), np.int64)
    np.put_along_axis(out, order, r2s, axis=1)
    return out


def rankdata_avg(a: np.ndarray, axis: int = 0) -> np.ndarray:
    """float64 average-tie ranks along ``axis`` (bottleneck.rankdata)."""
    a = np.asarray(a)
    moved = np.moveaxis(a, axis, -1)
    shp = moved.shape
    r = rank2_rows(moved.reshape(-1, shp[-1])).astype(np.float64) * 0.5
    return np.moveaxis(r.reshape(shp), -1, axis)


def _nanrank_cols(a: np.ndarray) -> np.ndarray:
    """bottleneck.nanrankdata(a, axis=0): NaNs ignored and returned as NaN
    (MetaNeighborUS.py:202, MetaNeighbor.py:232)."""
    a = np.asarray(a, np.float64)
    out = np.full(a.shape, np.nan)
    for j in range(a.shape[1]):
        ok = ~np.isnan(a[:, j])
        if ok.any():
            out[ok, j] = rank2_rows(a[ok, j]) * 0.5
    return out


# --------------------------------------------------------------------------
# normalisation  (utils.py:122-148)
# --------------------------------------------------------------------------
def normalize_cells(x: np.ndarray, ranked: bool = True) -> np.ndarray:
    """Per-cell rank, centre, L2-normalise.  Zero-variance cells -> NaN row
    (0/0), exactly as utils.py:143-147."""
    res = np.array(x, dtype=np.float64, copy=True)
    if ranked:
        res = rank2_rows(res).astype(np.float64) * 0.5
    res -= res.mean(axis=1)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        res /= np.sqrt(np.nansum(res ** 2, axis=1))[:, None]
    return res


# --------------------------------------------------------------------------
# labels  (utils.py:9-32, 78-119)
# --------------------------------------------------------------------------
def join_labels(s, c, replace_bar: bool = False) -> np.ndarray:
    """'$STUDY|$CELLTYPE' (utils.py:78-102).  replace_bar swaps '|' -> '.' in
    the study part only."""
    s = [str(v) for v in s]
    c = [str(v) for v in c]
    if replace_bar:
        s = [v.replace("|", ".") for v in s]
    return np.array([a + "|" + b for a, b in zip(s, c)], dtype=object)


def encode(vec):
    """Sorted unique values + integer codes: the column order of
    ``pd.get_dummies`` (utils.py:119) and of ``np.unique`` (utils.py:30)."""
    vec = np.asarray(vec, dtype=object)
    uniq = sorted(set(vec.tolist()))
    lut = {u: i for i, u in enumerate(uniq)}
    codes = np.fromiter((lut[v] for v in vec.tolist()), np.int64, len(vec))
    return np.array(uniq, dtype=object), codes


def _onehot(codes, n):
    m = np.zeros((codes.shape[0], n))
    m[np.arange(codes.shape[0]), codes] = 1.0
    return m


# --------------------------------------------------------------------------
# AUROC  (utils.py:151-191)
# --------------------------------------------------------------------------
def compute_aurocs(votes: np.ndarray, codes: np.ndarray, n_labels: int):
    """Rank-sum AUROC.  votes: (n_t, L); codes: label index of each test cell.
    Returns (n_labels, L) float64 with rows = test labels (utils.py:172-178)."""
    votes = np.asarray(votes, np.float64)
    n_t = votes.shape[0]
    ranks = rankdata_avg(votes, axis=0)
    # a NaN vote makes the whole column's ranks NaN (scipy.stats.rankdata's nan_policy="propagate", which
    # stands in for bottleneck.rankdata here -- oracle/ref_loader.py); only score_low_mem can produce one
    ranks[:, np.isnan(votes).any(axis=0)] = np.nan
    n_pos = np.bincount(codes, minlength=n_labels).astype(np.float64)
    n_neg = n_t - n_pos
    sum_pos = _onehot(codes, n_labels).T @ ranks
    with np.errstate(invalid="ignore", divide="ignore"):
        roc = sum_pos / n_pos[:, None]
        roc -= (n_pos[:, None] + 1) / 2
        roc /= n_neg[:, None]
    return roc, n_pos, n_neg


def compute_pvals(roc, n_pos, n_neg):
    """One-sided normal approximation of Mann-Whitney U (utils.py:180-188)."""
    from math import erfc, sqrt
    n_pos = np.asarray(n_pos, np.float64).reshape(-1, 1)
    n_neg = np.asarray(n_neg, np.float64).reshape(-1, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        u = roc * n_pos * n_neg
        z = np.abs(u - n_pos * n_neg / 2) / np.sqrt(n_pos * n_neg * (n_pos + n_neg + 1) / 12)
    sf = np.vectorize(lambda v: 0.5 * erfc(v / sqrt(2.0)) if v == v else np.nan, otypes=[float])
    return sf(z)


def _desc_order_nan_last(a: np.ndarray) -> np.ndarray:
    """Indices of ``a`` by descending value, NaN last, exact ties in ascending
    index order (deterministic rule; the reference's order on exact ties is
    implementation-defined: Series.sort_values(ascending=False), utils.py:227)."""
    idx = np.arange(a.shape[0])
    ok = ~np.isnan(a)
    good = idx[ok]
    order = good[np.argsort(-a[ok], kind="stable")]
    return np.concatenate([order, idx[~ok]])


def one_vs_best(votes: np.ndarray, codes: np.ndarray, aurocs: np.ndarray) -> np.ndarray:
    """compute_1v1_aurocs + find_top_candidates (utils.py:194-248).

    votes (n_t, L); codes (n_t,) test label of each cell; aurocs (L_t, L)."""
    votes = np.asarray(votes, np.float64)
    l_t, l = aurocs.shape
    out = np.full((l_t, l), np.nan)
    members = [np.flatnonzero(codes == c) for c in range(l_t)]
    for col in range(l):
        a = aurocs[:, col]
        if np.all(np.isnan(a)):
            continue
        cand = _desc_order_nan_last(a)[:5]
        best = cand[0]
        second = cand[1]
        score = 1.0
        v_best = votes[members[best], col]
        for cont in cand[1:]:
            v_cont = votes[members[cont], col]
            both = np.concatenate([v_best, v_cont])
            r = rank2_rows(both) * 0.5
            nb, nc = v_best.shape[0], v_cont.shape[0]
            auroc = (r[:nb].sum() / nb - (nb + 1) / 2) / nc
            if auroc < 0.5:
                second = best
                best = cont
                score = 1 - auroc
                v_best = v_cont
            elif auroc < score:
                score = auroc
                second = cont
        out[best, col] = score
        out[second, col] = 1 - score
    return out


# --------------------------------------------------------------------------
# fast / pretrained unsupervised path  (MetaNeighborUS.py:228-430)
# --------------------------------------------------------------------------
def _predict_and_score(x_norm, study, labels, centroids, n_per_cluster, col_labels,
                       col_study, ndn, one_vs_best_flag, compute_p):
    """MetaNeighborUS.py:310-387.  x_norm (N, G) normalised cells; study (N,)
    test study of each cell; labels (N,) joined label of each cell;
    centroids (G, L); col_study (L,) train study of each centroid column."""
    if ndn:
        st_uniq, st_codes = encode(col_study)
        study_matrix = _onehot(st_codes, len(st_uniq))            # L x S
        study_centroids = centroids @ study_matrix                # :349
        n_per_study = n_per_cluster @ study_matrix                # :350
    rows, blocks, pblocks = [], [], []
    study = np.asarray(study, dtype=object)
    for test_study in sorted(set(study.tolist())):                # :356 np.unique(S)
        is_test = study == test_study
        xt = x_norm[is_test]
        votes = xt @ centroids                                    # :361
        if ndn:
            votes = votes + n_per_cluster                         # :363
            deg = xt @ study_centroids + n_per_study              # :365-366
            votes = votes / deg[:, st_codes]                      # :368-371
        lab_uniq, lab_codes = encode(labels[is_test])
        roc, n_pos, n_neg = compute_aurocs(votes, lab_codes, len(lab_uniq))   # :375
        if compute_p:
            pblocks.append(compute_pvals(roc, n_pos, n_neg))
        if one_vs_best_flag:
            roc = one_vs_best(votes, lab_codes, roc)              # :382
        rows.extend(lab_uniq.tolist())
        blocks.append(roc)
    res = np.concatenate(blocks, axis=0)
    p = np.concatenate(pblocks, axis=0) if compute_p else None
    return res, np.array(rows, dtype=object), np.asarray(col_labels, dtype=object), p


def metaneighbor_us_fast(x, s, c, ndn=True, one_vs_best_flag=False, compute_p=False):
    """metaNeighborUS_fast (MetaNeighborUS.py:228-306).  Returns
    (table [test rows x train cols], row_labels, col_labels[, pvals])."""
    x_norm = normalize_cells(x)                                   # :270
    keep = ~np.any(np.isnan(x_norm), axis=1)                      # :273-277
    x_norm = x_norm[keep]
    s = np.asarray(s, dtype=object)[keep]
    c = np.asarray(c, dtype=object)[keep]
    labels = join_labels(s, c)                                    # :281
    lab_uniq, lab_codes = encode(labels)
    onehot = _onehot(lab_codes, len(lab_uniq))
    centroids = x_norm.T @ onehot                                 # :284
    n_per_cluster = onehot.sum(axis=0)                            # :288
    col_study = np.empty(len(lab_uniq), dtype=object)
    col_study[lab_codes] = s                                      # :339-340
    res, rows, cols, p = _predict_and_score(x_norm, s, labels, centroids, n_per_cluster,
                                            lab_uniq, col_study, ndn, one_vs_best_flag, compute_p)
    # :305  result = result[result.index]  -> columns reordered to the row order
    pos = {lab: i for i, lab in enumerate(cols.tolist())}
    perm = np.array([pos[r] for r in rows.tolist()])
    res = res[:, perm]
    if compute_p:
        return res, rows, rows.copy(), p[:, perm]
    return res, rows, rows.copy()


def train_model(x, s, c):
    """trainModel (trainModel.py:29-48) after HVG sub-setting.  Returns
    (model [(1+G) x L], column labels); row 0 = n_cells."""
    x = np.asarray(x)
    keep = np.asarray(x, np.float64).sum(axis=1) > 0              # :36-37
    x = x[keep]
    labels = join_labels(np.asarray(s, dtype=object)[keep], np.asarray(c, dtype=object)[keep])
    x_norm = normalize_cells(x)                                   # :39
    lab_uniq, lab_codes = encode(labels)
    onehot = _onehot(lab_codes, len(lab_uniq))
    cent = x_norm.T @ onehot                                      # :43
    n_cells = onehot.sum(axis=0)
    return np.vstack([n_cells[None, :], cent]), lab_uniq


def metaneighbor_us_from_trained(model, model_cols, x, s, c, ndn=True,
                                 one_vs_best_flag=False, compute_p=False):
    """MetaNeighborUS_from_trained (MetaNeighborUS.py:391-430); ``model`` is
    the (1+G) x L array already aligned to the query's gene order (:65-70)."""
    x_norm = normalize_cells(x)                                   # :410
    keep = ~np.any(np.isnan(x_norm), axis=1)                      # :411-412
    x_norm = x_norm[keep]
    s = np.asarray(s, dtype=object)[keep]
    c = np.asarray(c, dtype=object)[keep]
    labels = join_labels(s, c, replace_bar=True)                  # :417
    centroids = np.asarray(model[1:], np.float64)
    n_per_cluster = np.asarray(model[0], np.float64)
    col_study = np.array([str(v).split("|")[0] for v in model_cols], dtype=object)   # :341-343
    res, rows, cols, p = _predict_and_score(x_norm, s, labels, centroids, n_per_cluster,
                                            model_cols, col_study, ndn, one_vs_best_flag, compute_p)
    if compute_p:
        return res, rows, cols, p
    return res, rows, cols


# --------------------------------------------------------------------------
# full-network ("default") unsupervised path  (utils.py:35-75, MetaNeighborUS.py:145-224)
# --------------------------------------------------------------------------
def create_nw_spearman(x_cells_by_genes: np.ndarray) -> np.ndarray:
    """create_nw_spearman + rank (utils.py:35-75).  Input is cells x genes
    (the reference is handed the transpose and ranks along axis 0)."""
    r = rank2_rows(np.asarray(x_cells_by_genes)).astype(np.float64) * 0.5     # :69
    with np.errstate(invalid="ignore", divide="ignore"):
        nw = np.corrcoef(r, rowvar=True)                          # :71
    nw = np.atleast_2d(nw)
    np.fill_diagonal(nw, 1)                                       # :72
    finite = np.isfinite(nw)                                      # :44
    ranks = rank2_rows(nw[finite]).astype(np.float64) * 0.5       # :45
    ranks -= 1
    ranks /= ranks.max()
    nw[...] = 0.0                                                 # :50 nan_val=0
    nw[finite] = ranks
    np.fill_diagonal(nw, 1)                                       # :74
    return nw


def metaneighbor_us_default(x, s, c, ndn=True, compute_p=False):
    """metaNeighborUS_default + compute_aurocs_default (MetaNeighborUS.py:145-224).
    Returns table[train row, test col] (note the orientation)."""
    s = np.asarray(s, dtype=object)
    labels = join_labels(s, c)
    lab_uniq, lab_codes = encode(labels)                          # utils.py:30-31
    n_lab = len(lab_uniq)
    w = create_nw_spearman(x)
    sum_in = w @ _onehot(lab_codes, n_lab)                        # :165
    if ndn:
        sum_in = sum_in / w.sum(axis=0)[:, None]                  # :167-169
    out = np.full((n_lab, n_lab), np.nan)
    pout = np.full((n_lab, n_lab), np.nan)
    lab_study = np.empty(n_lab, dtype=object)
    lab_study[lab_codes] = s
    for study in sorted(set(s.tolist())):
        rows = np.flatnonzero(s == study)
        ranks = _nanrank_cols(sum_in[rows])                       # :202
        for ct in np.flatnonzero(lab_study == study):
            pos = lab_codes[rows] == ct
            n_p = float(pos.sum())
            nn = rows.shape[0] - n_p
            with np.errstate(invalid="ignore", divide="ignore"):
                roc = (ranks[pos].sum(axis=0) / n_p - (n_p + 1) / 2) / nn     # :209-212
            out[:, ct] = roc                                      # :213 column = test label
            if compute_p:
                pout[:, ct] = compute_pvals(roc[None, :], [n_p], [nn])[0]
    if compute_p:
        return out, lab_uniq, lab_uniq.copy(), pout
    return out, lab_uniq, lab_uniq.copy()


# --------------------------------------------------------------------------
# supervised MetaNeighbor  (MetaNeighbor.py:66-245)
# --------------------------------------------------------------------------
def _compute_votes(cand, voters, voter_id, ndn):
    """compute_votes (MetaNeighbor.py:151-176). cand (n_c, g); voters (g, n_v)."""
    votes = cand @ (voters @ voter_id)                            # :170
    if ndn:
        votes = votes + voter_id.sum(axis=0)                      # :172-173
        norm = cand @ voters.sum(axis=1) + voters.shape[1]        # :174
        votes = votes / norm[:, None]
    return votes


def score_low_mem(x, s, c, ndn=True):
    """score_low_mem (MetaNeighbor.py:106-148).  Returns (K,) AUROCs, K labels."""
    x = np.asarray(x, np.float64)
    keep = x.sum(axis=1) > 0                                      # :121
    x = x[keep]
    s = np.asarray(s, dtype=object)[keep]
    c = np.asarray(c, dtype=object)[keep]
    ct_uniq, ct_codes = encode(c)
    k = len(ct_uniq)
    onehot = _onehot(ct_codes, k)
    x_norm = normalize_cells(x)                                   # :130
    studies = sorted(set(s.tolist()))
    res = np.full((k, len(studies)), np.nan)
    for si, study in enumerate(studies):
        is_s = s == study
        votes = _compute_votes(x_norm[is_s], x_norm[~is_s].T, onehot[~is_s], ndn)
        present, codes_s = encode(ct_codes[is_s])                 # design_matrix(votes.index)
        roc, _, _ = compute_aurocs(votes, codes_s, len(present))
        for row, ct in enumerate(present.tolist()):               # :147 diag after reindex
            res[ct, si] = roc[row, ct]
    with np.errstate(invalid="ignore"), _quiet():
        return np.nanmean(res, axis=1), ct_uniq                   # :148


def score_default(x, s, c, ndn=True):
    """score_default (MetaNeighbor.py:179-245), means=True.  Returns (K,), labels."""
    s = np.asarray(s, dtype=object)
    nw = create_nw_spearman(x)                                    # :195
    nw = (nw + nw.T) / 2                                          # :196
    ct_uniq, ct_codes = encode(c)
    k = len(ct_uniq)
    onehot = _onehot(ct_codes, k)
    studies = sorted(set(s.tolist()))
    out = np.full((k, len(studies)), np.nan)
    deg = nw.sum(axis=0)                                          # :216
    for si, study in enumerate(studies):
        in_s = s == study
        voters = onehot * (~in_s)[:, None]                        # :203-211 hide test labels
        pred = nw @ voters                                        # :213
        if ndn:
            pred = pred / deg[:, None]                            # :217
        pred = np.abs(pred[in_s])                                 # :219-232 rows of study only
        ranks = rankdata_avg(pred, axis=0)
        for ct in range(k):
            pos = ct_codes[in_s] == ct
            n_p = float(pos.sum())
            n_n = float((~pos).sum())
            with np.errstate(invalid="ignore", divide="ignore"):
                out[ct, si] = (ranks[pos, ct].sum() / n_p - (n_p + 1) / 2) / n_n    # :234-238
    with np.errstate(invalid="ignore"), _quiet():
        return np.nanmean(out, axis=1), ct_uniq                   # :243


def metaneighbor_supervised(x, gene_names, s, c, genesets, geneset_index, geneset_names,
                            ndn=True, fast_version=False):
    """MetaNeighbor loop (MetaNeighbor.py:66-93).  genesets: (n_genes_gs, n_sets)
    0/1 array indexed by ``geneset_index``.  Returns (K x n_kept_sets), K labels,
    kept set names."""
    gene_names = np.asarray(gene_names, dtype=object)
    geneset_index = np.asarray(geneset_index, dtype=object)
    shared = np.array(sorted(set(gene_names.tolist()) & set(geneset_index.tolist())), dtype=object)  # :66
    gs_pos = {g: i for i, g in enumerate(geneset_index.tolist())}
    x_pos = {g: i for i, g in enumerate(gene_names.tolist())}
    gs = np.asarray(genesets)[[gs_pos[g] for g in shared.tolist()]]              # :71
    keep_sets = gs.sum(axis=0) > 0                                               # :72
    gs = gs[:, keep_sets].astype(bool)
    names = np.asarray(geneset_names, dtype=object)[keep_sets]
    expr = np.asarray(x)[:, [x_pos[g] for g in shared.tolist()]]                 # :81-83
    cols = []
    labels = None
    for j in range(gs.shape[1]):
        sub = expr[:, np.flatnonzero(gs[:, j])]                                  # :85
        if fast_version:
            r, labels = score_low_mem(sub, s, c, ndn)
        else:
            r, labels = score_default(sub, s, c, ndn)
        cols.append(r)
    return np.stack(cols, axis=1), labels, names


class _quiet:
    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore", RuntimeWarning)

    def __exit__(self, *a):
        return self._cm.__exit__(*a)


# ---------------------------------------------------------------------------
# variableGenes (reference pymn/variableGenes.py:7-77) -- SURVEY section 8f-2
# ---------------------------------------------------------------------------
def compute_var_genes(x: np.ndarray) -> np.ndarray:
    """variableGenes.py:22-37 for one dataset (cells x genes): per-gene median and population variance,
    median deciles (midpoint quantiles, right-closed digitize), variance quartiles inside each decile,
    keep the top quartile.  float64 throughout."""
    x = np.asarray(x, np.float64)
    median = np.median(x, axis=0)
    variance = np.var(x, axis=0)
    bins = np.quantile(median, q=np.linspace(0, 1, 11), method="midpoint")
    digits = np.digitize(median, bins, right=True)
    selected = np.zeros(x.shape[1], bool)
    for i in np.unique(digits):
        filt = digits == i
        var_tmp = variance[filt]
        bins_tmp = np.nanquantile(var_tmp, q=np.linspace(0, 1, 5))
        selected[filt] = np.digitize(var_tmp, bins_tmp) >= 4
    return selected


def variable_genes(x: np.ndarray, study) -> np.ndarray:
    """variableGenes.py:62-73: a gene must be selected in every study."""
    study = np.asarray(study, dtype=object)
    out = np.ones(np.asarray(x).shape[1], bool)
    for st in np.unique(study):
        out &= compute_var_genes(np.asarray(x)[study == st])
    return out
