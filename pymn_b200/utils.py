"""Host-side label logic of the hot path (mirrors reference pymn/utils.py:9-32,
78-119): joined ``study|celltype`` labels, their sorted order (the column order
``pd.get_dummies`` / ``np.unique`` give the reference), and the cell
permutation that makes studies and labels contiguous row ranges on the device.
Pure NumPy on integer codes -- no per-cell Python string work beyond one
``np.unique`` per obs column.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np


def join_labels(x, y, replace_bar=False):
    """'$STUDY|$CELLTYPE' (reference utils.py:78-102).  Vectorised; kept public
    because the reference re-exports it (pymn/__init__.py:18)."""
    x = np.asarray(x, dtype=object)
    y = np.asarray(y, dtype=object)
    if replace_bar:
        warnings.warn("Replacing any | with a . in study column values")
        x = np.array([str(v).replace("|", ".") for v in x], dtype=object)
    return np.array([str(a) + "|" + str(b) for a, b in zip(x, y)], dtype=object)


def _factorize_sorted(vec):
    """Sorted unique values (Python / code-point order, as np.unique on object
    or str arrays) and int64 codes.  ``vec``: numpy array, pandas Series or pandas array -- a pandas
    string column is factorised as it is (Arrow dictionary encoding), without a detour through
    200,000 Python objects."""
    import pandas as pd
    if isinstance(vec, pd.Series):
        vec = vec.values
    if not isinstance(vec, np.ndarray) and getattr(getattr(vec, "dtype", None), "name", "") == "category":
        vec = np.asarray(vec, dtype=object)
    # hash-factorise the N cells (O(N)), then sort only the handful of distinct values
    codes, uniq = pd.factorize(vec, sort=False)
    if (codes < 0).any():                      # missing values: the reference would stringify them ("nan")
        vec = np.array([str(v) for v in np.asarray(vec, dtype=object)], dtype=object)
        codes, uniq = pd.factorize(vec, sort=False)
    kind = np.asarray(vec[:1]).dtype.kind if len(vec) else "O"
    uniq = np.asarray(uniq, dtype=object) if kind in "OUST" else np.asarray(uniq)
    if kind in "OUST":
        names = [v if isinstance(v, str) else str(v) for v in uniq.tolist()]
        order = sorted(range(len(names)), key=lambda i: names[i])          # code-point order, as np.unique
    else:
        order = np.argsort(uniq, kind="stable").tolist()                      # numeric order, as np.unique
        names = [str(v) for v in uniq.tolist()]
    remap = np.empty(len(order), np.int64)
    remap[np.asarray(order, dtype=np.int64)] = np.arange(len(order))
    return np.array([names[i] for i in order], dtype=object), remap[np.asarray(codes, dtype=np.int64)]


def _stable_argsort_small(codes: np.ndarray, n_values: int) -> np.ndarray:
    """Stable argsort of non-negative integer codes < n_values; 16-bit keys take NumPy's O(N) radix sort."""
    if n_values <= 32767:
        return np.argsort(codes.astype(np.int16), kind="stable")
    return np.argsort(codes, kind="stable")


@dataclass
class LabelLayout:
    """Device row layout: cells sorted by (study, joined label).

    ``row_labels``  joined labels in table row order (for each study in sorted
                    order, that study's labels sorted -- MetaNeighborUS.py:356,375)
    ``perm``        int32[N]: device row i holds input cell perm[i]
    ``cell_label``  int32[N]: row-order label id of device row i
    ``cell_study``  int32[N]: study index of device row i
    ``seg_off``     int32[S+1]: study t owns device rows [seg_off[t], seg_off[t+1])
    ``lab_row_off`` int32[L+1]: label c owns device rows [lab_row_off[c], lab_row_off[c+1])
    ``seg_lab_off`` int32[S+1]: study t owns labels [seg_lab_off[t], seg_lab_off[t+1])
    ``label_study`` int32[L]: study index of each label
    ``sorted_pos``  int64[L]: position of each row-order label in the globally
                    sorted label list (np.unique(study_ct), utils.py:30)
    """
    studies: np.ndarray
    row_labels: np.ndarray
    perm: np.ndarray
    cell_label: np.ndarray
    cell_study: np.ndarray
    seg_off: np.ndarray
    lab_row_off: np.ndarray
    seg_lab_off: np.ndarray
    label_study: np.ndarray
    sorted_pos: np.ndarray


def _factorize_two(a, b):
    """The two obs columns are factorised side by side (the Arrow / pandas hash kernels release the GIL)."""
    n = len(a)
    if n < 100_000:
        return _factorize_sorted(a), _factorize_sorted(b)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=2) as ex:
        fa, fb = ex.submit(_factorize_sorted, a), ex.submit(_factorize_sorted, b)
        return fa.result(), fb.result()


def label_layout(study, ctype, replace_bar=False) -> LabelLayout:
    (st_uniq, st_codes), (ct_uniq, ct_codes) = _factorize_two(study, ctype)
    n_ct = len(ct_uniq)
    # (32-bit codes when they fit: the O(N) passes below are memory-bound on the host)
    if len(st_uniq) * n_ct < (1 << 31):
        st_codes, ct_codes = st_codes.astype(np.int32), ct_codes.astype(np.int32)
    pair = st_codes * n_ct + ct_codes
    # distinct (study, type) pairs, ascending, and each cell's index into them (O(N): the code space is small)
    n_codes = len(st_uniq) * n_ct
    if n_codes <= max(4 * len(pair), 1 << 20):
        present = np.bincount(pair, minlength=n_codes) > 0
        pair_uniq = np.flatnonzero(present)
        pair_codes = (np.cumsum(present) - 1)[pair]
    else:
        pair_uniq, pair_codes = np.unique(pair, return_inverse=True)
    st_of = pair_uniq // n_ct
    st_names = [s.replace("|", ".") for s in st_uniq.tolist()] if replace_bar else st_uniq.tolist()
    if replace_bar:
        warnings.warn("Replacing any | with a . in study column values")
    names = np.array([st_names[s] + "|" + ct_uniq[c] for s, c in zip(st_of.tolist(), (pair_uniq % n_ct).tolist())],
                     dtype=object)
    # row order: study (sorted) major, joined label (sorted as a string) minor
    order = sorted(range(len(names)), key=lambda i: (int(st_of[i]), names[i]))
    order = np.asarray(order, dtype=np.int64)
    rank_of_pair = np.empty(len(names), np.int32)
    rank_of_pair[order] = np.arange(len(names), dtype=np.int32)
    cell_label = rank_of_pair[pair_codes]
    perm = _stable_argsort_small(cell_label, len(names))
    row_labels = names[order]
    label_study = st_of[order].astype(np.int32)
    n_lab = len(names)
    counts = np.bincount(cell_label, minlength=n_lab)
    lab_row_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_st = len(st_uniq)
    seg_lab_off = np.searchsorted(label_study, np.arange(n_st + 1)).astype(np.int32)
    seg_off = lab_row_off[seg_lab_off].astype(np.int32)
    glob = sorted(range(n_lab), key=lambda i: row_labels[i])
    sorted_pos = np.empty(n_lab, np.int64)
    sorted_pos[np.asarray(glob, dtype=np.int64)] = np.arange(n_lab)
    cell_label_sorted = cell_label[perm].astype(np.int32)
    return LabelLayout(studies=st_uniq, row_labels=row_labels, perm=perm.astype(np.int32),
                       cell_label=cell_label_sorted,
                       cell_study=label_study[cell_label_sorted].astype(np.int32),
                       seg_off=seg_off, lab_row_off=lab_row_off, seg_lab_off=seg_lab_off,
                       label_study=label_study, sorted_pos=sorted_pos)


def joint_layout(study, ctype):
    """Layout for the supervised paths: label id = study * K + type over ALL
    S*K combinations (empty ranges allowed), cells sorted by that id.
    Returns (LabelLayout, cell-type names [K])."""
    st_uniq, st_codes = _factorize_sorted(study)
    ct_uniq, ct_codes = _factorize_sorted(ctype)
    S, K = len(st_uniq), len(ct_uniq)
    joint = st_codes * K + ct_codes
    perm = _stable_argsort_small(joint, S * K)
    counts = np.bincount(joint, minlength=S * K)
    lab_row_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    seg_lab_off = (np.arange(S + 1) * K).astype(np.int32)
    seg_off = lab_row_off[seg_lab_off].astype(np.int32)
    names = np.array([f"{s}|{c}" for s in st_uniq.tolist() for c in ct_uniq.tolist()], dtype=object)
    lay = LabelLayout(studies=st_uniq, row_labels=names, perm=perm.astype(np.int32),
                      cell_label=joint[perm].astype(np.int32), cell_study=st_codes[perm].astype(np.int32),
                      seg_off=seg_off, lab_row_off=lab_row_off, seg_lab_off=seg_lab_off,
                      label_study=np.repeat(np.arange(S), K).astype(np.int32),
                      sorted_pos=np.arange(S * K, dtype=np.int64))
    return lay, ct_uniq


def select_genes(adata, var_genes):
    """``adata[:, var_genes]`` -- except that selecting every gene in its own order returns ``adata`` itself:
    AnnData (and AnnDataLite) materialise a view's ``X`` as a copy, 1.6 GB of host traffic at config 2 for
    an identity selection."""
    names = adata.var_names
    try:
        if len(var_genes) == len(names) and (np.asarray(var_genes) == np.asarray(names)).all():
            return adata
    except Exception:
        pass
    return adata[:, var_genes]


def prepare_matrix(X):
    """What the device pipelines ingest: a scipy CSR matrix as it is (AnnData's native layout; only
    the stored values are uploaded), anything else as a dense C-contiguous float32 array."""
    from scipy import sparse
    if sparse.issparse(X):
        return X.tocsr()
    return np.ascontiguousarray(np.asarray(X), dtype=np.float32)


def to_dense_f32(X) -> np.ndarray:
    """Dense C-contiguous float32 cells x genes (the reference densifies at
    utils.py:136-139; CSR ingestion on device is SURVEY section 8f 'next')."""
    from scipy import sparse
    if sparse.issparse(X):
        X = X.toarray()
    return np.ascontiguousarray(np.asarray(X), dtype=np.float32)
