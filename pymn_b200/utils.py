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


_POOL = None
# Encode the two obs columns on two threads (1) or one after the other (0, the default).  The helper thread saves
# ~1.5 ms of a 200,000-cell call -- and, on busy hosts, costs a 30-60 ms stall every few calls when the sleeping
# thread is woken late (measured back to back on one box: 104-106 ms with one 167 ms call, against 106-108 ms flat).
LABEL_HELPER = __import__("os").environ.get("MN_LABEL_HELPER", "0") != "0"


def _pool():
    """One persistent helper thread for the host-side label encoding (Arrow's hash kernels release the GIL;
    starting a thread per call costs as much as the work)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="pymn_b200-labels")
    return _POOL


def _arrow_flat(vec):
    """The Arrow array behind a pandas string array (one chunk), or None when ``vec`` is not Arrow-backed / holds nulls."""
    arr = getattr(vec, "_pa_array", None)
    if arr is None:
        return None
    try:
        import pyarrow as pa
        if arr.null_count or not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type)):
            return None
        return arr.combine_chunks() if isinstance(arr, pa.ChunkedArray) else arr
    except Exception:
        return None


def _arrow_encode(flat):
    """Sorted distinct values (code-point order, as np.unique) and int32 codes of an Arrow string array: Arrow's own
    dictionary encoding (the hash kernel runs without the GIL), then a remap of the handful of distinct values."""
    enc = flat.dictionary_encode()
    codes = np.asarray(enc.indices.to_numpy(zero_copy_only=False), dtype=np.int32)
    first_seen = enc.dictionary.to_pylist()
    names = sorted(first_seen)
    pos = {v: i for i, v in enumerate(names)}
    remap = np.asarray([pos[v] for v in first_seen], dtype=np.int32)
    return np.array(names, dtype=object), (remap[codes] if len(first_seen) else codes)


def _factorize_sorted(vec):
    """Sorted unique values (Python / code-point order, as np.unique on object
    or str arrays) and integer codes.  ``vec``: numpy array, pandas Series or pandas array -- a pandas
    string column is factorised as it is (Arrow dictionary encoding), without a detour through
    200,000 Python objects."""
    import pandas as pd
    if isinstance(vec, pd.Series):
        vec = vec.values
    if not isinstance(vec, np.ndarray) and getattr(getattr(vec, "dtype", None), "name", "") == "category":
        vec = np.asarray(vec, dtype=object)
    flat = _arrow_flat(vec)
    if flat is not None:
        return _arrow_encode(flat)
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
    """Stable argsort of non-negative integer codes < n_values; 8- and 16-bit keys take NumPy's O(N) radix sort
    (one / two counting passes)."""
    if n_values <= 256:
        return np.argsort(codes.astype(np.uint8), kind="stable")
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
    """The two obs columns are factorised side by side: one on the persistent helper thread, one on the calling
    thread (Arrow's hash kernels and pandas' hash tables release the GIL).  One hand-off per call, on purpose: with
    eight slice tasks on four threads the interpreter's 5 ms GIL switch interval showed up as sporadic 50 ms calls."""
    import pandas as pd
    a = a.values if isinstance(a, pd.Series) else a
    b = b.values if isinstance(b, pd.Series) else b
    if len(a) < 50_000 or not LABEL_HELPER:
        return _factorize_sorted(a), _factorize_sorted(b)
    fb = _pool().submit(_factorize_sorted, b)
    ra = _factorize_sorted(a)
    return ra, fb.result()


def label_layout(study, ctype, replace_bar=False) -> LabelLayout:
    (st_uniq, st_codes), (ct_uniq, ct_codes) = _factorize_two(study, ctype)
    n_ct = len(ct_uniq)
    n_codes = len(st_uniq) * n_ct
    # (32-bit codes when they fit: the O(N) passes below are memory-bound on the host)
    if n_codes < (1 << 31):
        st_codes, ct_codes = st_codes.astype(np.int32, copy=False), ct_codes.astype(np.int32, copy=False)
    pair = st_codes * n_ct + ct_codes
    # distinct (study, type) pairs, ascending, and how many cells each holds (O(N): the code space is small)
    small_space = n_codes <= max(4 * len(pair), 1 << 20)
    if small_space:
        pair_cnt = np.bincount(pair, minlength=n_codes)
        pair_uniq = np.flatnonzero(pair_cnt)
        cnt_uniq = pair_cnt[pair_uniq]
    else:
        pair_uniq, pair_codes, cnt_uniq = np.unique(pair, return_inverse=True, return_counts=True)
    st_of = pair_uniq // n_ct
    st_names = [s.replace("|", ".") for s in st_uniq.tolist()] if replace_bar else st_uniq.tolist()
    if replace_bar:
        warnings.warn("Replacing any | with a . in study column values")
    names = np.array([st_names[s] + "|" + ct_uniq[c] for s, c in zip(st_of.tolist(), (pair_uniq % n_ct).tolist())],
                     dtype=object)
    # row order: study (sorted) major, joined label (sorted as a string) minor
    order = sorted(range(len(names)), key=lambda i: (int(st_of[i]), names[i]))
    order = np.asarray(order, dtype=np.int64)
    n_lab = len(names)
    rank_of_pair = np.empty(n_lab, np.int32)
    rank_of_pair[order] = np.arange(n_lab, dtype=np.int32)
    if small_space:
        # one gather through a table over the code space: pair code -> row-order label id
        lut = np.zeros(n_codes, np.int32)
        lut[pair_uniq] = rank_of_pair
        cell_label = lut[pair]
    else:
        cell_label = rank_of_pair[pair_codes]
    perm = _stable_argsort_small(cell_label, n_lab)
    row_labels = names[order]
    label_study = st_of[order].astype(np.int32)
    counts = cnt_uniq[order]                   # cells per row-order label: no further pass over the cells
    lab_row_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_st = len(st_uniq)
    seg_lab_off = np.searchsorted(label_study, np.arange(n_st + 1)).astype(np.int32)
    seg_off = lab_row_off[seg_lab_off].astype(np.int32)
    glob = sorted(range(n_lab), key=lambda i: row_labels[i])
    sorted_pos = np.empty(n_lab, np.int64)
    sorted_pos[np.asarray(glob, dtype=np.int64)] = np.arange(n_lab)
    # the sorted order is label 0 repeated counts[0] times, label 1 ...: written sequentially, not gathered
    cell_label_sorted = np.repeat(np.arange(n_lab, dtype=np.int32), counts)
    return LabelLayout(studies=st_uniq, row_labels=row_labels, perm=perm.astype(np.int32),
                       cell_label=cell_label_sorted,
                       cell_study=np.repeat(label_study, counts),
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


class gc_paused:
    """Pause Python's cyclic garbage collector for the duration of one API call.  The call allocates a few
    thousand short-lived containers (label factorisation, DataFrames); when that trips a full collection, the
    collector walks the CALLER's whole heap -- with a few million label strings alive (the obs columns of the
    very AnnData being scored) that is 10-20 ms, a fifth of the GPU work of a 200,000-cell call.  Nothing in the
    call creates reference cycles worth collecting early; the previous state is restored on exit."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        if self._was:
            gc.disable()
        return self

    def __exit__(self, *exc):
        if self._was:
            import gc
            gc.enable()
        return False


def with_gc_paused(fn):
    """Decorator form of gc_paused for the public entry points (signature and docstring preserved)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with gc_paused():
            return fn(*args, **kwargs)
    return wrapper


MAX_GENES = 16384


def check_gene_count(n_genes: int, what: str = "selected genes"):
    """The rank kernels index genes with 16 bits: more than 16,384 genes per cell cannot be ranked in one call.
    The reference has no such limit (it is 100-1000x slower there, but it runs): fail early and clearly."""
    if int(n_genes) > MAX_GENES:
        raise ValueError(f"pymn_b200 ranks at most {MAX_GENES} {what} per cell (got {int(n_genes)}); "
                         "select highly variable genes first (pymn_b200.variableGenes)")


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
