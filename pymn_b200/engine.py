"""Device pipelines of the MetaNeighbor hot path.

Thin Python over the C ABI (``_lib.call``): allocates torch CUDA buffers,
enqueues the kernels of ``libpymn_b200.so`` on the current stream and returns
small result tables to the host.  No arithmetic of the hot path happens here
(torch is plumbing: memory, streams, H2D/D2H); the few torch ops used are
index bookkeeping on O(N) int vectors and concatenation of O(L) vectors.

Pipelines (reference call stacks in SURVEY.md section 3):
  us_fast          metaNeighborUS_fast + predict_and_score   MetaNeighborUS.py:228-387
  us_pretrained    MetaNeighborUS_from_trained               MetaNeighborUS.py:391-430
  us_default       metaNeighborUS_default                    MetaNeighborUS.py:145-224, utils.py:35-75
  train_model      trainModel                                trainModel.py:29-48
  supervised_*     score_default / score_low_mem             MetaNeighbor.py:106-245

Multi-GPU (one process per GPU, torch.distributed / NCCL): the centroid pipelines shard the CELLS
(cell_shard: integer all-reduce of fixed-point centroid sums, vote GEMM storing into peer GPUs' receive buffers,
all-gather of the table's column slices), the full-network pipeline shards the row tiles of its two passes
(integer all-reduces), the supervised pipeline shards the gene sets.  Uploads: AsyncUpload (pinned host matrix),
StagedUpload (pageable, worker-thread staging), ShardedUpload (each rank its slice of the cells).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .utils import LabelLayout

import os as _os

# CDF resolution of the global rank (bins over rho in [-1, 1]); MN_NBINS is a tuning knob
DEFAULT_NBINS = int(_os.environ.get("MN_NBINS", "16384"))


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


_STAGE = {}                      # device index -> (pinned uint8 [2, STAGE_BYTES], [event, event])
STAGE_BYTES = 64 << 20


def _upload_pageable(a: np.ndarray, dev) -> torch.Tensor:
    """A large PAGEABLE host array -> device through two reusable pinned 64 MB buffers: the host copies chunk
    i + 1 into one buffer while the DMA engine drains the other.  Measured for a pageable 1.6 GB matrix on the B200 box: 37 ms
    (43 GB/s; 97 ms on the first call, which allocates the staging area) against 840 ms for a cold
    `pin_memory()` of the whole array (65 ms once torch's pinned cache holds a block that large) and 142 ms
    for a plain pageable `.to()`."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _STAGE:
        _STAGE[idx] = (torch.empty((2, STAGE_BYTES), dtype=torch.uint8, pin_memory=True),
                       [torch.cuda.Event(), torch.cuda.Event()])
    stage, evs = _STAGE[idx]
    src = torch.from_numpy(a.reshape(-1).view(np.uint8))
    out = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=dev)
    dst = out.view(-1).view(torch.uint8)
    n = src.shape[0]
    for i, off in enumerate(range(0, n, STAGE_BYTES)):
        b = i & 1
        evs[b].synchronize()                     # the copy that last used this buffer has drained
        m = min(STAGE_BYTES, n - off)
        stage[b, :m].copy_(src[off:off + m])           # (torch's host copy is multi-threaded for large blocks)
        dst[off:off + m].copy_(stage[b, :m], non_blocking=True)
        evs[b].record()
    return out


def to_device(x, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (pinned arrays go as they are, large pageable ones through
    a double-buffered pinned staging area)."""
    dev = _dev()
    if isinstance(x, torch.Tensor):
        t = x.to(dev, non_blocking=True)
    else:
        a = np.ascontiguousarray(x)
        t = torch.from_numpy(a)
        if a.nbytes >= (8 << 20) and not t.is_pinned():
            t = _upload_pageable(a, dev)
        else:
            # Anything above 32 KB goes through pinned memory (torch's caching host allocator: ~0.1 ms for 1 MB).  A
            # PAGEABLE copy of more than 64 KB is staged piecewise by the driver, and its later pieces queue in the copy
            # engine behind whatever was issued meanwhile -- e.g. the 1.2 GB train of the streamed matrix upload: the
            # label arrays then arrive 20 ms late and the first kernel with them (tools/h2d_probe.py: a stream
            # waiting for the 3rd of 12 queued copies resumes at 5.5 ms without, at 21.7 ms with such a copy).
            if a.nbytes >= (32 << 10):
                t = t.pin_memory()
            t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _i32(a):
    return to_device(np.asarray(a, dtype=np.int32))


@dataclass
class CsrMatrix:
    """Expression matrix in CSR form on the device (scipy.sparse.csr_matrix layout)."""
    indptr: torch.Tensor       # int64 [n_rows + 1]
    indices: torch.Tensor      # int32 [nnz]
    data: torch.Tensor         # float32 [nnz]
    shape: tuple


@dataclass
class CscMatrix:
    """Expression matrix in CSC form on the device (scipy.sparse.csc_matrix layout): column access for the
    gene-set slices of the supervised path and the per-gene statistics of variableGenes."""
    indptr: torch.Tensor       # int64 [n_cols + 1]
    indices: torch.Tensor      # int32 [nnz] row ids
    data: torch.Tensor         # float32 [nnz]
    shape: tuple


def to_device_csc(X) -> CscMatrix:
    """scipy.sparse matrix -> CscMatrix (only the stored values cross PCIe)."""
    csc = X.tocsc()
    if not csc.has_canonical_format:
        csc = csc.copy()
        csc.sum_duplicates()
    return CscMatrix(to_device(np.asarray(csc.indptr, dtype=np.int64)), to_device(np.asarray(csc.indices, dtype=np.int32)),
                     to_device(np.asarray(csc.data, dtype=np.float32)), tuple(csc.shape))


def csc_columns_dense(csc: CscMatrix, cols: torch.Tensor = None, col0: int = 0, n_cols: int = None) -> torch.Tensor:
    """Dense float32 [N, g] tile of the columns ``cols`` (int32 device tensor) or [col0, col0 + n_cols)."""
    g = int(cols.numel()) if cols is not None else int(n_cols)
    N = int(csc.shape[0])
    ld = (g + 3) // 4 * 4                       # rows stay 16-byte aligned for K1's 128-bit loads
    out = torch.empty((N, ld), dtype=torch.float32, device=csc.data.device)
    _lib.call("mn_csc_gather_dense", csc.indptr, csc.indices, csc.data, cols, int(col0), g, N, out, ld)
    return out[:, :g] if ld != g else out


def to_device_matrix(X):
    """Dense array / tensor -> float32 CUDA tensor; scipy.sparse matrix -> CsrMatrix (no densification:
    only the stored values cross PCIe)."""
    from scipy import sparse
    if isinstance(X, CsrMatrix):
        return X
    if sparse.issparse(X):
        csr = X.tocsr()
        if not csr.has_canonical_format:
            csr = csr.copy()
            csr.sum_duplicates()
        return CsrMatrix(to_device(np.asarray(csr.indptr, dtype=np.int64)),
                         to_device(np.asarray(csr.indices, dtype=np.int32)),
                         to_device(np.asarray(csr.data, dtype=np.float32)), tuple(csr.shape))
    return to_device(X, torch.float32)


@dataclass
class RankedCells:
    """K1 output.  Full-network / test form: fp16 planes d_hi (+ d_lo beyond 2,048 genes).  Centroid-path form
    (``rank_normalize(..., limbs=True)``): no planes, ``_limbs`` = (a1, a0, ldk, inv_norm64) int8 limb planes."""
    d_hi: torch.Tensor | None
    d_lo: torch.Tensor | None
    ldd: int
    inv_norm: torch.Tensor
    flags: torch.Tensor
    M: int
    G: int
    xnorm: torch.Tensor | None = None
    rank2: torch.Tensor | None = None
    _limbs: tuple | None = None

    @property
    def device(self):
        return self.inv_norm.device


def alloc_ranked(M: int, G: int, dev) -> RankedCells:
    """Uninitialised K1 outputs for M cells x G genes (filled by rank_normalize(..., out=, out_row0=))."""
    ldd = (G + 63) // 64 * 64
    return RankedCells(torch.empty((M, ldd), dtype=torch.float16, device=dev),
                       torch.empty((M, ldd), dtype=torch.float16, device=dev) if G > 2048 else None, ldd,
                       torch.empty((M,), dtype=torch.float32, device=dev),
                       torch.empty((M,), dtype=torch.uint8, device=dev), M, G)


def rank_normalize(X: torch.Tensor, row_index=None, col_index=None, drop_mode=0,
                   want_xnorm=False, want_rank2=False, out: RankedCells = None, out_row0: int = 0,
                   limbs: bool = False) -> RankedCells:
    """K1.  X: float32 CUDA [N, Gx] row-major, or a CsrMatrix.  With ``out`` the M = len(row_index) cells are
    written into rows [out_row0, out_row0 + M) of preallocated outputs (the streamed pipeline ranks the
    cells chunk by chunk as their rows arrive).  ``limbs``: write the int8 limb planes + double scale of the
    centroid paths instead of the fp16 planes (mn_rank_normalize_limbs)."""
    if limbs:
        assert out is None and not want_xnorm and not want_rank2
        is_csr = isinstance(X, CsrMatrix)
        dev = X.data.device if is_csr else X.device
        assert not (is_csr and col_index is not None), "column sub-setting of a CSR matrix is not supported"
        M = int(row_index.numel()) if row_index is not None else int(X.shape[0])
        G = int(col_index.numel()) if col_index is not None else int(X.shape[1])
        ldk = (G + 127) // 128 * 128
        a1 = torch.empty((M, ldk), dtype=torch.int8, device=dev)
        a0 = torch.empty((M, ldk), dtype=torch.int8, device=dev)
        inv64 = torch.empty((M,), dtype=torch.float64, device=dev)
        inv_norm = torch.empty((M,), dtype=torch.float32, device=dev)
        flags = torch.empty((M,), dtype=torch.uint8, device=dev)
        if M > 0 and is_csr:
            _lib.call("mn_rank_normalize_csr_limbs", X.indptr, X.indices, X.data, row_index, M, G, a1, a0, ldk, inv64,
                      inv_norm, flags, int(drop_mode))
        elif M > 0:
            _lib.call("mn_rank_normalize_limbs", X, int(X.stride(0)), row_index, col_index, M, G, a1, a0, ldk, inv64,
                      inv_norm, flags, int(drop_mode))
        return RankedCells(None, None, (G + 63) // 64 * 64, inv_norm, flags, M, G, _limbs=(a1, a0, ldk, inv64))
    if out is not None:
        assert row_index is not None and col_index is None and not want_xnorm and not want_rank2
        M = int(row_index.numel())
        r0, r1 = int(out_row0), int(out_row0) + M
        if M > 0 and isinstance(X, CsrMatrix):
            _lib.call("mn_rank_normalize_csr", X.indptr, X.indices, X.data, row_index, M, out.G, out.d_hi[r0:r1],
                      out.d_lo[r0:r1] if out.d_lo is not None else None, out.ldd, out.inv_norm[r0:r1],
                      out.flags[r0:r1], None, None, int(drop_mode))
        elif M > 0:
            _lib.call("mn_rank_normalize", X, int(X.stride(0)), row_index, None, M, out.G, out.d_hi[r0:r1],
                      out.d_lo[r0:r1] if out.d_lo is not None else None, out.ldd, out.inv_norm[r0:r1],
                      out.flags[r0:r1], None, None, int(drop_mode))
        return out
    is_csr = isinstance(X, CsrMatrix)
    if is_csr:
        assert col_index is None, "column sub-setting of a CSR matrix is not supported"
        dev = X.data.device
    else:
        assert X.is_cuda and X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1     # (rows may be pitched)
        dev = X.device
    M = int(row_index.numel()) if row_index is not None else int(X.shape[0])
    G = int(col_index.numel()) if col_index is not None else int(X.shape[1])
    ldd = (G + 63) // 64 * 64
    d_hi = torch.empty((M, ldd), dtype=torch.float16, device=dev)
    d_lo = torch.empty((M, ldd), dtype=torch.float16, device=dev) if G > 2048 else None
    inv_norm = torch.empty((M,), dtype=torch.float32, device=dev)
    flags = torch.empty((M,), dtype=torch.uint8, device=dev)
    xnorm = torch.empty((M, G), dtype=torch.float32, device=dev) if want_xnorm else None
    rank2 = torch.empty((M, G), dtype=torch.int32, device=dev) if want_rank2 else None
    if M > 0 and is_csr:
        _lib.call("mn_rank_normalize_csr", X.indptr, X.indices, X.data, row_index, M, G, d_hi, d_lo, ldd,
                  inv_norm, flags, xnorm, rank2, int(drop_mode))
    elif M > 0:
        _lib.call("mn_rank_normalize", X, int(X.stride(0)), row_index, col_index, M, G, d_hi, d_lo, ldd,
                  inv_norm, flags, xnorm, rank2, int(drop_mode))
    return RankedCells(d_hi, d_lo, ldd, inv_norm, flags, M, G, xnorm, rank2)


def gene_stats(X, perm: np.ndarray, seg_off: np.ndarray):
    """Per-study per-gene median and variance (variableGenes.py:22-27).  X: dense host / device matrix
    [N, G] or a scipy.sparse matrix; perm int32[N] sorts the cells by study, seg_off int32[S+1] are the study ranges of that order.
    Returns (median float64 [S, G], variance float64 [S, G]) on the host."""
    from scipy import sparse
    S = int(len(seg_off) - 1)
    if sparse.issparse(X):
        # sparse AnnData.X: CSC on the device, 512 genes at a time expanded into a dense tile (never the whole matrix)
        csc = to_device_csc(X)
        N, G = csc.shape
        dev = csc.data.device
        med = torch.empty((S, G), dtype=torch.float64, device=dev)
        var = torch.empty((S, G), dtype=torch.float64, device=dev)
        perm_d, seg_d = _i32(perm), _i32(seg_off)
        step = 512
        for c0 in range(0, G, step):
            g = min(step, G - c0)
            tile = csc_columns_dense(csc, None, c0, g)
            m_c = torch.empty((S, g), dtype=torch.float64, device=dev)
            v_c = torch.empty((S, g), dtype=torch.float64, device=dev)
            _lib.call("mn_gene_stats", tile, int(tile.stride(0)), perm_d, seg_d, S, g, m_c, v_c)
            med[:, c0:c0 + g] = m_c
            var[:, c0:c0 + g] = v_c
        return med.cpu().numpy(), var.cpu().numpy()
    Xd = to_device(X, torch.float32)
    G = int(Xd.shape[1])
    med = torch.empty((S, G), dtype=torch.float64, device=Xd.device)
    var = torch.empty((S, G), dtype=torch.float64, device=Xd.device)
    _lib.call("mn_gene_stats", Xd, int(Xd.stride(0)), _i32(perm), _i32(seg_off), S, G, med, var)
    return med.cpu().numpy(), var.cpu().numpy()


def centroids_fixed(rc: RankedCells, row_off: torch.Tensor, L: int, extra_rows: int = 0, max_label_rows: int = 0):
    """K2, fixed-point form.  Returns (Cq int64 [L + extra_rows, ldd] = sum over the label's kept cells of
    x* rounded to multiples of 2^-38, nq int32 [L + extra_rows] kept cells): exact integer sums, so partial
    results of different GPUs add (all-reduce) to bit-identical centroids for every world size."""
    dev = rc.device
    Cq = torch.zeros((L + extra_rows, rc.ldd), dtype=torch.int64, device=dev)
    nq = torch.zeros((L + extra_rows,), dtype=torch.int32, device=dev)
    if rc.M > 0 and L > 0:
        a1, a0, ldk, inv64 = cell_limbs(rc)
        mlr = int(max_label_rows) if max_label_rows else rc.M
        if rc.d_hi is None:
            _lib.call("mn_centroids_fixed_i8", a1, a0, ldk, inv64, row_off, L, rc.G, mlr, Cq, rc.ldd, nq)
        else:
            _lib.call("mn_centroids_fixed", rc.d_hi, rc.d_lo, rc.ldd, inv64, row_off, L, rc.G, mlr, Cq, rc.ldd, nq)
    return Cq, nq


def group_sum_fixed_into(Cq, nq, L, group: torch.Tensor, n_groups: int):
    """Study centroids / study sizes appended as rows L.. of Cq / nq (integer sums of the label rows)."""
    ld = int(Cq.stride(0))
    _lib.call("mn_group_sum_fixed", Cq, ld, L, ld, group, n_groups, Cq[L:], nq, nq[L:])


def centroids_finish(Cq, nq):
    """(C float32 [R, ld], n_valid float32 [R]) of fixed-point centroid sums."""
    R, ld = int(Cq.shape[0]), int(Cq.stride(0))
    C = torch.empty((R, ld), dtype=torch.float32, device=Cq.device)
    n_valid = torch.empty((R,), dtype=torch.float32, device=Cq.device)
    _lib.call("mn_centroids_finish", Cq, nq, R, ld, C, n_valid)
    return C, n_valid


def centroids(rc: RankedCells, row_off: torch.Tensor, L: int, extra_rows: int = 0, max_label_rows: int = 0):
    """K2.  Returns (C float32 [L + extra_rows, ldd] (extra rows zero), n_valid float32 [L + extra_rows])."""
    return centroids_finish(*centroids_fixed(rc, row_off, L, extra_rows, max_label_rows))


def group_sum_into(C, n_valid, L, group: torch.Tensor, n_groups: int, width: int):
    """Study centroids / study sizes appended as rows L.. of C / n_valid (float form: pretrained models)."""
    ld = int(C.stride(0))
    _lib.call("mn_group_sum", C, ld, L, width, group, n_groups, C[L:])
    _lib.call("mn_group_sum", n_valid, 1, L, 1, group, n_groups, n_valid[L:])


def split_f16(C: torch.Tensor, G: int):
    R, ld = int(C.shape[0]), int(C.stride(0))
    hi = torch.empty((R, ld), dtype=torch.float16, device=C.device)
    lo = torch.empty((R, ld), dtype=torch.float16, device=C.device)
    inv_scale = torch.empty((R,), dtype=torch.float32, device=C.device)
    _lib.call("mn_split_f16", C, ld, R, G, hi, lo, ld, inv_scale)
    return hi, lo, inv_scale


def cell_limbs(rc: RankedCells):
    """int8 limb planes of the ranked cells (d = 128 a1 + a0) and their double-precision scale, cached on ``rc``."""
    if getattr(rc, "_limbs", None) is None:
        dev = rc.device
        ldk = (rc.G + 127) // 128 * 128
        a1 = torch.empty((rc.M, ldk), dtype=torch.int8, device=dev)
        a0 = torch.empty((rc.M, ldk), dtype=torch.int8, device=dev)
        inv64 = torch.empty((rc.M,), dtype=torch.float64, device=dev)
        if rc.M > 0:
            _lib.call("mn_limb_split_cells", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, rc.M, rc.G, a1, a0, ldk, inv64)
        rc._limbs = (a1, a0, ldk, inv64)
    return rc._limbs


def centroid_limbs(rc: RankedCells, C: torch.Tensor, Nn: int):
    """The Nn centroid rows of C as four stacked int8 digit planes: (limbs [4 * rows_pad, ldk], rows_pad, inv_scale [Nn])."""
    dev = rc.device
    ldk = cell_limbs(rc)[2]
    rows_pad = (Nn + 127) // 128 * 128
    limbs = torch.zeros((4 * rows_pad, ldk), dtype=torch.int8, device=dev)
    inv_scale = torch.empty((Nn,), dtype=torch.float32, device=dev)
    _lib.call("mn_limb_split_centroids", C, int(C.stride(0)), Nn, rc.G, limbs, ldk, rows_pad, inv_scale, None)
    return limbs, rows_pad, inv_scale


def votes_exact_launch(rc: RankedCells, cl, addend, Nn: int, out: torch.Tensor, r0: int = 0, r1: int = None, den=None,
                       col_den=None, col_scale=None):
    """One launch of the exact vote GEMM for the cells [r0, r1) of ``rc`` against the centroid digits ``cl``
    (centroid_limbs): column-major votes into ``out`` [Nn, ldo >= r1 - r0] (float32, or float64 = the raw values)."""
    a1, a0, ldk, inv64 = cell_limbs(rc)
    r1 = rc.M if r1 is None else int(r1)
    if r1 <= r0:
        return
    limbs, rows_pad, inv_scale = cl
    as_double = out.dtype == torch.float64
    _lib.call("mn_votes_fast_i8", a1[r0:r1], a0[r0:r1], ldk, inv64[r0:r1], r1 - r0, limbs, ldk, rows_pad, inv_scale, addend,
              Nn, rc.G, None if as_double else out, out if as_double else None, int(out.stride(0)),
              den[:, r0:r1] if den is not None else None, int(den.stride(0)) if den is not None else 0, col_den, col_scale)


def votes_exact(rc: RankedCells, C: torch.Tensor, addend, Nn: int, den=None, col_den=None, col_scale=None,
                as_double=False, ldo: int = 0):
    """K3, exact form (the product path): votes of all cells for the Nn centroid rows of C (float32 [>= Nn, ld])
    through integer limb products on the int8 tensor cores; the epilogue works in double:
    val = x.c + addend;  with ``den`` (double [S, M]) val = val / den[col_den[n]] * col_scale[n] - 1, otherwise
    val -= col_scale[n].  Returns [Nn, max(M, ldo)] column-major, float32 (rounded once) or float64 (``as_double``)."""
    dev = rc.device
    ldo = max(int(ldo), rc.M)          # (a pitch > M leaves zeroed padding)
    out = (torch.zeros if ldo > rc.M else torch.empty)((Nn, ldo), dtype=torch.float64 if as_double else torch.float32,
                                                       device=dev)
    votes_exact_launch(rc, centroid_limbs(rc, C, Nn), addend, Nn, out, 0, rc.M, den, col_den, col_scale)
    return out


def votes_den64(rc: RankedCells, C: torch.Tensor, addend, S: int, cuda_cores: bool = False):
    """Node degree of every cell against the S study-centroid rows of C, in double: [S, M].  Exact integer dot
    products either way: all eight digit products on the int8 tensor cores (mn_votes_den_i8, the product path) or
    64-bit integer FMAs on the CUDA cores (mn_votes_den64, ``cuda_cores`` / MN_DEN_CUDA_CORES=1: the A/B reference)."""
    dev = rc.device
    a1, a0, ldk, inv64 = cell_limbs(rc)
    inv_scale = torch.empty((S,), dtype=torch.float32, device=dev)
    out = torch.empty((S, max(rc.M, 1)), dtype=torch.float64, device=dev)
    if cuda_cores or _os.environ.get("MN_DEN_CUDA_CORES"):
        q32 = torch.empty((S, ldk), dtype=torch.int32, device=dev)
        _lib.call("mn_limb_split_centroids", C, int(C.stride(0)), S, rc.G, None, ldk, S, inv_scale, q32)
        if rc.M > 0:
            _lib.call("mn_votes_den64", a1, a0, ldk, inv64, rc.M, rc.G, q32, inv_scale, addend, S, out, int(out.stride(0)))
        return out
    rows_pad = (S + 127) // 128 * 128
    limbs = torch.zeros((4 * rows_pad, ldk), dtype=torch.int8, device=dev)
    _lib.call("mn_limb_split_centroids", C, int(C.stride(0)), S, rc.G, limbs, ldk, rows_pad, inv_scale, None)
    if rc.M > 0:
        _lib.call("mn_votes_den_i8", a1, a0, ldk, inv64, rc.M, limbs, ldk, rows_pad, inv_scale, addend, S, rc.G, out,
                  int(out.stride(0)))
    return out


def votes_fast(rc: RankedCells, b_hi, b_lo, inv_scale, addend, Nn: int):
    """K3.  Returns float32 [Nn, M] (column-major votes: one row per train column)."""
    out = torch.empty((Nn, rc.M), dtype=torch.float32, device=rc.d_hi.device)
    _lib.call("mn_votes_fast", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, rc.M, b_hi, b_lo, int(b_hi.stride(0)), inv_scale,
              addend, Nn, rc.G, out, rc.M)
    return out


def auroc(votes: torch.Tensor, ncols: int, denom, col_denom, seg_off: torch.Tensor, n_seg: int,
          cell_label: torch.Tensor, n_labels: int, out: torch.Tensor = None):
    """K6.  votes float32 [>=ncols, M].  Returns (auroc float64 [n_labels, ncols], n_pos int32 [n_labels])."""
    M = int(votes.shape[1])
    dev = votes.device
    if out is None:
        out = torch.empty((n_labels, ncols), dtype=torch.float64, device=dev)
    n_pos = torch.empty((n_labels,), dtype=torch.int32, device=dev)
    wbytes = _lib.auroc_workspace_bytes(ncols, M, n_seg, n_labels)
    ws = torch.empty((wbytes,), dtype=torch.uint8, device=dev)
    _lib.call("mn_auroc", votes, int(votes.stride(0)), ncols, denom, col_denom, seg_off, n_seg, cell_label,
              n_labels, M, out, n_pos, ws, wbytes)
    return out, n_pos


def mask_dropped(cell_label: torch.Tensor, inv_norm: torch.Tensor, flags=None, poison=None) -> torch.Tensor:
    """Test labels with the dropped cells (inv_norm == 0) set to -1; ``poison`` (uint8 [1]) is raised when a cell
    is constant but non-zero (see mn_mask_dropped)."""
    out = torch.empty_like(cell_label)
    _lib.call("mn_mask_dropped", cell_label, inv_norm, flags, int(cell_label.numel()), out, poison)
    return out


def one_vs_best(votes, ncols, denom, col_denom, seg_lab_off, n_seg, lab_row_off, cell_label, n_labels, auc):
    """K7.  Returns float64 [n_labels, ncols], NaN except two cells per (study, column)."""
    out = torch.full((n_labels, ncols), float("nan"), dtype=torch.float64, device=votes.device)
    _lib.call("mn_one_vs_best", votes, int(votes.stride(0)), ncols, denom, col_denom, seg_lab_off, n_seg,
              lab_row_off, cell_label, n_labels, auc, out)
    return out


# ---------------------------------------------------------------------------
# pipelines
# ---------------------------------------------------------------------------
REF_BLOCK = 4096        # input cells whose mean vote centres a column (see _column_reference)


@dataclass
class CellShard:
    """The cells one rank owns in the centroid paths (cells are the data-parallel unit).

    Rank r owns the contiguous block [lo, hi) of INPUT rows (what it uploads through the public API); ``rows``
    lists them in (study, label)-sorted order -- a subsequence of the global device row order ``lay.perm`` --
    and ``lab_row_off`` are the label ranges of that local order.  ``gidx[p]`` is where global device row p
    sits in the concatenation of the ranks' (zero-padded, ``per`` cells each) local orders: a gather through it
    puts exchanged per-cell arrays into global device row order."""
    rank: int
    world: int
    N: int
    per: int
    lo: int
    hi: int
    rows: np.ndarray
    lab_row_off: np.ndarray
    max_label_rows: int
    gidx: np.ndarray

    def on_device(self, name: str, make):
        """Device copy of a host array derived from this shard, uploaded once per device (the bench and repeated API
        calls on one layout reuse it)."""
        cache = self.__dict__.setdefault("_dev", {})
        key = (name, torch.cuda.current_device())
        if key not in cache:
            cache[key] = make()
        return cache[key]


def cell_shard(lay: LabelLayout, rank: int, world: int) -> CellShard:
    """The cell shard of (rank, world) for a layout, computed once per layout (a few O(N) host passes)."""
    cache = lay.__dict__.setdefault("_shards", {})
    if (rank, world) not in cache:
        cache[(rank, world)] = _cell_shard(lay, rank, world)
    return cache[(rank, world)]


def _cell_shard(lay: LabelLayout, rank: int, world: int) -> CellShard:
    perm = np.asarray(lay.perm)
    N, L = int(perm.shape[0]), len(lay.row_labels)
    if world == 1:
        counts = np.diff(np.asarray(lay.lab_row_off))
        return CellShard(0, 1, N, N, 0, N, perm, np.asarray(lay.lab_row_off, np.int32),
                         int(counts.max()) if L else 0, None)
    lo, hi, per = shard_rows(N, rank, world)
    mine = (perm >= lo) & (perm < hi)
    counts = np.bincount(np.asarray(lay.cell_label)[mine], minlength=L)
    src = perm // per
    lpos = np.empty(N, np.int64)
    for r in range(world):
        idx = np.flatnonzero(src == r)
        lpos[idx] = np.arange(idx.shape[0])
    return CellShard(rank, world, N, per, lo, hi, perm[mine], np.concatenate([[0], np.cumsum(counts)]).astype(np.int32),
                     int(counts.max()) if L else 0, src.astype(np.int64) * per + lpos)


def lay_dev(lay: LabelLayout, name: str) -> torch.Tensor:
    """int32 device copy of one of the layout's index arrays, uploaded once per layout and device."""
    cache = lay.__dict__.setdefault("_dev", {})
    key = (name, torch.cuda.current_device())
    if key not in cache:
        cache[key] = _i32(getattr(lay, name))
    return cache[key]


def ctypes_addr(arr) -> int:
    """Address of a ctypes array (a host pointer argument of the C ABI)."""
    import ctypes
    return ctypes.addressof(arr)


def _gather_cells(local: torch.Tensor, sh: CellShard, gidx: torch.Tensor) -> torch.Tensor:
    """Per-cell array of the own cells (local order) -> the same for all N cells in global device row order."""
    if sh.world == 1:
        return local
    return gather_rows(local, sh.per, sh.world).index_select(0, gidx)


def _all_reduce_sum(*tensors):
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def _device_matrix(X):
    """(expression matrix on the device, input row id of its first row) for whatever the caller handed in:
    a ShardedUpload (the rank holds only its block of input rows), an AsyncUpload, a host / device array or CSR."""
    if isinstance(X, ShardedUpload):
        return X.X_local, X.lo
    if isinstance(X, AsyncUpload):
        return X.wait_all(), 0
    return to_device_matrix(X), 0


def local_ranked_cells(Xd, base: int, sh: CellShard, drop_mode=0):
    """K1 over this rank's cells in local label-sorted order (``Xd`` row 0 = input row ``base``)."""
    rows = sh.on_device(f"rows-{base}", lambda: _i32(sh.rows - base))
    return rank_normalize(Xd, row_index=rows, drop_mode=drop_mode, limbs=True)


def _reference_votes(tot: torch.Tensor, C: torch.Tensor, n_rows: int) -> torch.Tensor:
    """mu[n] = (mean cell) . C[n] in double, from a fixed-point cell sum ``tot`` (int64 [ldd + 1]: sums, then the count)."""
    ldd = int(C.stride(0))
    mean_cell = tot[:ldd].double() * (2.0 ** -38) / tot[ldd].clamp(min=1).double()
    return (C[:n_rows].double() * mean_cell).sum(dim=1)


def _column_reference(Xd, base: int, sh: CellShard, C: torch.Tensor, n_rows: int) -> torch.Tensor:
    """mu[n] = mean over the first min(N, REF_BLOCK) INPUT cells of x* . C[n] (double [n_rows]): the value a
    column's votes are centred on before their one float32 rounding (any constant is a monotone map of the
    column, to which K6 / K7 are invariant; a typical vote keeps the rounding at ~1e-9 of the vote where the
    cells crowd).  The block's cell sum is a fixed-point integer sum (all-reduced), so every world size
    centres on the same value."""
    dev = C.device
    n_ref = min(sh.N, REF_BLOCK)
    b_lo, b_hi = sh.lo, min(sh.hi, n_ref)
    ldd = int(C.stride(0))
    if b_hi > b_lo:
        rb = rank_normalize(Xd, row_index=torch.arange(b_lo - base, b_hi - base, dtype=torch.int32, device=dev),
                            drop_mode=0, limbs=True)
        m = b_hi - b_lo
        nb = (m + 31) // 32
        off = torch.clamp(torch.arange(0, nb + 1, dtype=torch.int32, device=dev) * 32, max=m)
        Sq, nq = centroids_fixed(rb, off, nb, max_label_rows=32)
        tot = torch.cat([Sq.sum(dim=0), nq.sum().to(torch.int64).reshape(1)])
    else:
        tot = torch.zeros((ldd + 1,), dtype=torch.int64, device=dev)
    if sh.world > 1:
        _all_reduce_sum(tot)
    return _reference_votes(tot, C, n_rows)


def _score_against_centroids(Xd, base, rc: RankedCells, sh: CellShard, lay: LabelLayout, C, n_valid, n_train, train_study,
                             n_train_studies, ndn: bool, ovb: bool, timer=None, ref_tot=None):
    """predict_and_score (MetaNeighborUS.py:310-387) for all test studies at once.
    C rows [0, n_train) = cluster centroids, rows [n_train, n_train + n_train_studies) = study centroids
    (identical on every rank); ``rc`` = this rank's ranked cells (local order of ``sh``).

    Multi-GPU (one process per GPU): cells are the unit of the vote GEMM -- every rank computes the votes of ITS
    cells for all train columns -- and train columns are the unit of the AUROC sort and the one-vs-best
    contest, so one all-to-all turns [all columns, own cells] into [own columns, all cells] (column j lives on
    rank j % world), and one all-gather assembles the L_test x L_train table.  Votes are exact integer dot
    products scaled in double, so every world size produces bit-identical tables."""
    dev = rc.device
    rank, world = sh.rank, sh.world
    n_lab = len(lay.row_labels)
    n_seg = len(lay.studies)
    t = timer
    n_max = (n_train + world - 1) // world
    n_loc = len(range(rank, n_train, world))
    gidx = sh.on_device("gidx", lambda: to_device(sh.gidx)) if world > 1 else None
    flags = _gather_cells(rc.flags, sh, gidx)
    cell_label = mask_dropped(lay_dev(lay, "cell_label"), _gather_cells(rc.inv_norm, sh, gidx))
    seg_off = lay_dev(lay, "seg_off")
    if t: t.start("votes_gemm")
    # Exact integer dot products; the node-degree ratio is taken in double inside the epilogue and centred on a
    # per-column reference so that its one float32 rounding resolves ~1e-9 of the ratio where the cells crowd
    # (rounding numerator and denominator to float32 first costs 3-4 rank flips per AUROC cell at 5,000 cells
    # per study).
    n_b = n_train + (n_train_studies if ndn else 0)
    if t: t.start("votes_ref")
    # (the centring reference: the mean of ALL cells when the caller has their exact fixed-point sum -- the label
    #  centroids add up to it --, else the mean of the first REF_BLOCK input cells)
    mu = _reference_votes(ref_tot, C, n_b) if ref_tot is not None else _column_reference(Xd, base, sh, C, n_b)
    if t: t.stop("votes_ref")
    if world > 1:
        # columns in destination order: rank d's columns d, d + world, ... as one block of n_max rows (padding
        # rows repeat column 0 and are dropped after the exchange)
        order = np.zeros(world * n_max, np.int64)
        for d in range(world):
            cols_d = np.arange(d, n_train, world)
            order[d * n_max:d * n_max + len(cols_d)] = cols_d
        order = to_device(order)
    else:
        order = None
    pick = (lambda v: v.index_select(0, order).contiguous()) if order is not None else (lambda v: v[:n_train].contiguous())
    n_cols = world * n_max if world > 1 else n_train
    if ndn:
        addend = n_valid[:n_b].contiguous()
        num_ref = mu[:n_train] + addend[:n_train].double()
        den_ref = (mu[n_train:] + addend[n_train:].double()).index_select(0, train_study.long())
        ref = num_ref / den_ref
        scale = torch.where(torch.isfinite(ref) & (ref > 0), 1.0 / ref, torch.ones_like(ref))
        if t: t.start("votes_den")
        den64 = votes_den64(rc, C[n_train:n_b], addend[n_train:n_b].contiguous(), n_train_studies)
        if t: t.stop("votes_den")
        kw = dict(den=den64, col_den=pick(train_study), col_scale=pick(scale))
        add_cols = pick(addend)
    else:
        kw = dict(col_scale=pick(mu))
        add_cols = None
    cl = centroid_limbs(rc, pick(C), n_cols)
    peer = peer_receive_buffers(n_cols * sh.per * 4, dev, rank, world) if (world > 1 and VOTES_P2P and world <= 8) else None
    if world == 1:
        out = torch.empty((n_cols, rc.M), dtype=torch.float32, device=dev)
        votes_exact_launch(rc, cl, add_cols, n_cols, out, 0, rc.M, **kw)
    elif peer is not None:
        # GEMM fused with the exchange: the epilogue stores every column's votes straight into its owner's receive
        # buffer over NVLink (mn_votes_fast_i8_peer) -- no collective moves the votes.  Two tiny all-reduces fence
        # the ranks: before the launch (everybody has finished reading its buffer from the previous call), after it
        # (all stores have landed).
        recv, table = peer
        fence = torch.zeros((1,), dtype=torch.int32, device=dev)
        _all_reduce_sum(fence)
        a1, a0, ldk, inv64 = cell_limbs(rc)
        limbs, rows_pad, inv_scale = cl
        if rc.M > 0:
            den = kw.get("den")
            _lib.call("mn_votes_fast_i8_peer", a1, a0, ldk, inv64, rc.M, limbs, ldk, rows_pad, inv_scale, add_cols, n_cols,
                      rc.G, ctypes_addr(table), world, rank, n_max, sh.per, den, int(den.stride(0)) if den is not None else 0,
                      kw.get("col_den"), kw.get("col_scale"))
        _all_reduce_sum(fence)
        if t: t.start("votes_exchange_tail")
        out = exchange_votes_end(recv[:world * n_max * sh.per].view(1, world * n_max, sh.per), [], sh, n_max, gidx)
        if t: t.stop("votes_exchange_tail")
    else:
        # The own cells go through the GEMM in Q row chunks; the all-to-all of chunk q runs (on NCCL's stream) while
        # the tensor cores work on chunk q + 1, so only the last chunk's exchange is exposed.
        Q = votes_chunks(sh.per)
        clen = ((sh.per + Q - 1) // Q + 127) // 128 * 128
        send = torch.zeros((Q, n_cols, clen), dtype=torch.float32, device=dev)
        for q in range(Q):
            votes_exact_launch(rc, cl, add_cols, n_cols, send[q], q * clen, min(rc.M, (q + 1) * clen), **kw)
            if q == 0:
                recv, works = exchange_votes_begin(send, q)
            else:
                exchange_votes_begin(send, q, recv, works)
        if t: t.start("votes_exchange_tail")
        out = exchange_votes_end(recv, works, sh, n_max, gidx)
        if t: t.stop("votes_exchange_tail")
        del send
    if t: t.stop("votes_gemm")
    if n_loc > 0:
        if t: t.start("auroc")
        auc, n_pos = auroc(out, n_loc, None, None, seg_off, n_seg, cell_label, n_lab)
        if t: t.stop("auroc")
        res = auc
        if ovb:
            if t: t.start("one_vs_best")
            res = one_vs_best(out, n_loc, None, None, lay_dev(lay, "seg_lab_off"), n_seg, lay_dev(lay, "lab_row_off"),
                              cell_label, n_lab, auc)
            if t: t.stop("one_vs_best")
    else:
        auc = res = torch.empty((n_lab, 0), dtype=torch.float64, device=dev)
        n_pos = torch.zeros((n_lab,), dtype=torch.int32, device=dev)
    if world == 1:
        return res, auc, n_pos, flags
    return (_gather_columns(res, n_train, rank, world), _gather_columns(auc, n_train, rank, world),
            _allreduce_max(n_pos), flags)


# ---- peer receive buffers of the fused vote GEMM + exchange (mn_votes_fast_i8_peer) ---------------------------------
VOTES_P2P = _os.environ.get("MN_VOTES_P2P", "1") != "0"
_PEER = {}          # device index -> dict(nbytes, local uint8 tensor, peers [uint8 tensors of every rank], table)


def peer_receive_buffers(nbytes: int, dev, rank: int, world: int):
    """A receive buffer of >= nbytes on every rank, each mapped into every other rank's address space (CUDA IPC through
    torch's storage sharing; one process per GPU on one node, NVLink / NVSwitch peer access).  Returns (local float32
    view, host array of the `world` device pointers) or None when the mapping cannot be set up (the caller falls
    back to NCCL's all-to-all).  Collective: every rank must call it with the same nbytes.  The buffers are cached
    for the life of the process and only ever grow."""
    import ctypes
    import torch.distributed as dist
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ent = _PEER.get(idx)
    if ent is not None and ent.get("failed"):
        return None
    if ent is None or ent["nbytes"] < nbytes:
        try:
            size = max(int(nbytes), 1 << 20)
            size = (size + (64 << 20) - 1) // (64 << 20) * (64 << 20)          # grow in 64 MB steps
            local = torch.empty((size,), dtype=torch.uint8, device=dev)
            info = local.untyped_storage()._share_cuda_()
            infos = [None] * world
            dist.all_gather_object(infos, (rank, info))
            peers = []
            for r in range(world):
                if r == rank:
                    peers.append(local)
                    continue
                # opened with THIS rank's device current (not the exporter's, which torch's tensor sharing would use):
                # cudaIpcOpenMemHandle maps the memory into the current device's address space and enables peer
                # access lazily -- what a kernel of this device needs in order to store into it (as NCCL does)
                info_r = (idx,) + tuple(infos[r][1][1:])
                st = torch.UntypedStorage._new_shared_cuda(*info_r)
                t = torch.empty((0,), dtype=torch.uint8, device=dev)
                t.set_(st, 0, (size,), (1,))
                peers.append(t)
            # touching every peer once through torch enables peer access between the two devices for this context
            probe = torch.zeros((16,), dtype=torch.uint8, device=dev)
            for r in range(world):
                if r != rank:
                    probe.copy_(peers[r][:16])
            torch.cuda.synchronize()
            ok = torch.ones((1,), dtype=torch.int32, device=dev)
        except Exception as e:                                 # noqa: BLE001 -- any failure = no peer path on this box
            ok = torch.zeros((1,), dtype=torch.int32, device=dev)
            peers, local, size = None, None, 0
            _PEER_ERR["last"] = f"{type(e).__name__}: {e}"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)              # all ranks take the same path
        if int(ok.item()) == 0:
            _PEER[idx] = {"failed": True}
            return None
        table = (ctypes.c_void_p * world)(*[int(p.data_ptr()) for p in peers])
        ent = _PEER[idx] = dict(nbytes=size, local=local, peers=peers, table=table)
    n32 = ent["nbytes"] // 4
    return ent["local"].view(torch.float32)[:n32], ent["table"]


_PEER_ERR = {}


def release_peer_buffers():
    """Drop the peer mappings and the local receive buffer (collective in spirit: call it on every rank before the
    process group goes away, so that no rank exits while another still maps its buffer)."""
    for ent in _PEER.values():
        ent.pop("peers", None)
        ent.pop("local", None)
    _PEER.clear()
    torch.cuda.empty_cache()
VOTES_CHUNKS = int(_os.environ.get("MN_VOTES_CHUNKS", "8"))


def votes_chunks(per: int) -> int:
    """Row chunks of the cell-sharded vote GEMM (each followed by its own all-to-all): 8 when the chunks stay large
    (only the last chunk's exchange is exposed: NCCL's all-to-all moves ~110 GB/s per rank here)."""
    return max(1, min(VOTES_CHUNKS, per // 4096))


def exchange_votes_begin(send: torch.Tensor, q: int, recv: torch.Tensor = None, works: list = None):
    """Start the all-to-all of row chunk q.  ``send`` [Q, world * n_max, clen]: votes of this rank's cells (local order,
    chunk q = cells [q clen, (q + 1) clen), zero-padded) for ALL train columns, destination rank d's columns in rows
    [d n_max, (d + 1) n_max).  Asynchronous: the collective waits for the work queued on the current stream (the
    GEMM of this chunk) and runs on the process group's own stream.  Returns (recv, works)."""
    import torch.distributed as dist
    if recv is None:
        recv, works = torch.empty_like(send), []
    works.append(dist.all_to_all_single(recv[q].view(-1), send[q].view(-1), async_op=True))
    return recv, works


def exchange_votes_end(recv: torch.Tensor, works: list, sh: CellShard, n_max: int, gidx: torch.Tensor) -> torch.Tensor:
    """Wait for the chunk exchanges and put the cells into global device row order: [n_max, N] (this rank's train
    columns for the cells of all ranks).  ``gidx`` addresses rank-major blocks of ``sh.per`` cells; the chunked
    buffers hold blocks of Q * clen >= per cells per rank."""
    for w in works:
        w.wait()
    Q, clen, world = int(recv.shape[0]), int(recv.shape[2]), sh.world
    per_pad = Q * clen
    if recv.is_cuda:
        # one pass: cell p sits at chunk q, source rank src, position i -> offset ((q world + src) n_max) clen + i
        src, lpos = gidx // sh.per, gidx % sh.per
        off = ((lpos // clen) * world + src) * (n_max * clen) + lpos % clen
        out = torch.empty((n_max, sh.N), dtype=recv.dtype, device=recv.device)
        _lib.call("mn_votes_unpack", recv, off, sh.N, n_max, clen, out, sh.N)
        return out
    # (host tensors -- the gloo tests of this bookkeeping -- take the same route with torch index operations)
    gv = (gidx // sh.per) * per_pad + gidx % sh.per if per_pad != sh.per else gidx
    # [chunk, source rank, column, cell in chunk] -> [column, source rank, chunk, cell in chunk] = [column, padded cell]
    return recv.view(Q, world, n_max, clen).permute(2, 1, 0, 3).reshape(n_max, world * per_pad).index_select(1, gv)


def exchange_votes(out: torch.Tensor, sh: CellShard, n_max: int, gidx: torch.Tensor) -> torch.Tensor:
    """The votes exchange of the cell-sharded centroid path in one piece: ``out`` [world * n_max, per] -> [n_max, N]."""
    recv, works = exchange_votes_begin(out.unsqueeze(0), 0)
    return exchange_votes_end(recv, works, sh, n_max, gidx)


def _gather_columns(local: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the per-rank column slices (column j lives on rank j % world) into [rows, n_total]."""
    import torch.distributed as dist
    n_max = (n_total + world - 1) // world
    buf = torch.full((local.shape[0], n_max), float("nan"), dtype=local.dtype, device=local.device)
    buf[:, :local.shape[1]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.empty((local.shape[0], n_total), dtype=local.dtype, device=local.device)
    for r in range(world):
        n_r = len(range(r, n_total, world))
        out[:, r::world] = parts[r][:, :n_r]
    return out


def _allreduce_max(t: torch.Tensor) -> torch.Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def ranked_cells(X, lay: LabelLayout, drop_mode=0) -> RankedCells:
    """K1 over ALL cells, in label-sorted order (the full-network and supervised paths), however the matrix
    arrives: host / device array or CSR (upload, then rank), an AsyncUpload (wait for the chunks), or a
    ShardedUpload (multi-GPU: rank the own slice, all-gather the planes)."""
    if isinstance(X, ShardedUpload):
        return ranked_cells_sharded(X, lay, drop_mode)
    Xd = X.wait_all() if isinstance(X, AsyncUpload) else to_device_matrix(X)
    return rank_normalize(Xd, row_index=_i32(lay.perm), drop_mode=drop_mode)


def _host_tables(res, auc, n_pos, flags):
    return dict(table=res.cpu().numpy(), auroc=auc.cpu().numpy(), n_pos=n_pos.cpu().numpy(), flags=flags.cpu().numpy())


def _label_centroids(rc: RankedCells, sh: CellShard, L: int, extra_rows: int = 0):
    """Fixed-point label centroids over the cells of ALL ranks: local partial sums + one integer all-reduce."""
    Cq, nq = centroids_fixed(rc, sh.on_device("lab_row_off", lambda: _i32(sh.lab_row_off)), L, extra_rows, sh.max_label_rows)
    if sh.world > 1:
        _all_reduce_sum(Cq, nq)
    return Cq, nq


def us_fast_device(X, lay: LabelLayout, ndn=True, ovb=False, timer=None):
    """metaNeighborUS_fast with every result left on the device: (table, auroc, n_pos, flags).
    With torch.distributed initialised on > 1 ranks the cells are sharded (cell_shard)."""
    t = timer
    rank, world, _ = dist_shard()
    sh = cell_shard(lay, rank, world)
    if t: t.start("rank_normalize")
    Xd, base = _device_matrix(X)
    rc = local_ranked_cells(Xd, base, sh, 0)
    if t: t.stop("rank_normalize")
    L, S = len(lay.row_labels), len(lay.studies)
    if t: t.start("centroids")
    Cq, nq = _label_centroids(rc, sh, L, extra_rows=S)
    train_study = lay_dev(lay, "label_study")
    if ndn:
        group_sum_fixed_into(Cq, nq, L, train_study, S)
    C, n_valid = centroids_finish(Cq, nq)
    # the label centroids add up to the exact (integer, all-reduced) sum of all kept cells: the centring reference
    ref_tot = torch.cat([Cq[:L].sum(dim=0), nq[:L].sum().to(torch.int64).reshape(1)])
    if t: t.stop("centroids")
    return _score_against_centroids(Xd, base, rc, sh, lay, C, n_valid, L, train_study, S, ndn, ovb, timer, ref_tot=ref_tot)


def us_fast(X, lay: LabelLayout, ndn=True, ovb=False):
    """metaNeighborUS_fast.  X: numpy / torch float32 [N, G] (input cell order), CSR, or an upload in flight
    (start_upload).  Returns dict(table [L, L] rows = test labels, cols = train labels (both row order),
    auroc (one-vs-rest table), n_pos [L], flags uint8 [N] in device row order)."""
    return _host_tables(*us_fast_device(X, lay, ndn, ovb))


def us_pretrained_device(X, lay: LabelLayout, model_centroids, model_n, model_study,
                         n_model_studies: int, ndn=True, ovb=False, timer=None):
    """MetaNeighborUS_from_trained with every result left on the device.  The model arrays may be host
    arrays or device tensors: centroids float [L_train, G] (rows = model columns), n [L_train], study int [L_train]."""
    t = timer
    rank, world, _ = dist_shard()
    sh = cell_shard(lay, rank, world)
    if t: t.start("rank_normalize")
    Xd, base = _device_matrix(X)
    rc = local_ranked_cells(Xd, base, sh, 0)
    if t: t.stop("rank_normalize")
    if t: t.start("centroids")
    Lt = int(model_centroids.shape[0])
    dev = rc.device
    C = torch.zeros((Lt + n_model_studies, rc.ldd), dtype=torch.float32, device=dev)
    C[:Lt, :rc.G] = (model_centroids if isinstance(model_centroids, torch.Tensor)
                     else to_device(np.asarray(model_centroids, np.float32)))
    n_valid = torch.zeros((Lt + n_model_studies,), dtype=torch.float32, device=dev)
    n_valid[:Lt] = model_n if isinstance(model_n, torch.Tensor) else to_device(np.asarray(model_n, np.float32))
    train_study = model_study if isinstance(model_study, torch.Tensor) else _i32(model_study)
    if ndn:
        group_sum_into(C, n_valid, Lt, train_study, n_model_studies, rc.G)
    if t: t.stop("centroids")
    return _score_against_centroids(Xd, base, rc, sh, lay, C, n_valid, Lt, train_study, n_model_studies, ndn, ovb, timer)


def us_pretrained(X, lay: LabelLayout, model_centroids: np.ndarray, model_n: np.ndarray, model_study: np.ndarray,
                  n_model_studies: int, ndn=True, ovb=False):
    """MetaNeighborUS_from_trained.  model_centroids float [L_train, G] (rows = model
    columns), model_n [L_train], model_study int [L_train]."""
    return _host_tables(*us_pretrained_device(X, lay, model_centroids, model_n, model_study, n_model_studies, ndn, ovb))


def train_model(X, lay: LabelLayout):
    """trainModel: returns dict(centroids float32 [L, G], n_cells [L], flags [N] device row order).
    Multi-GPU: every rank sums its own cells, one integer all-reduce."""
    rank, world, _ = dist_shard()
    sh = cell_shard(lay, rank, world)
    rc = local_ranked_cells(*_device_matrix(X), sh, 1)
    L = len(lay.row_labels)
    C, n_valid = centroids_finish(*_label_centroids(rc, sh, L))
    flags = _gather_cells(rc.flags, sh, sh.on_device("gidx", lambda: to_device(sh.gidx)) if world > 1 else None)
    return dict(centroids=C[:, :rc.G].cpu().numpy(), n_cells=n_valid.cpu().numpy(), flags=flags.cpu().numpy())


LAST_SPILL = {}         # what the last rho_spill_buffer call decided (bench.py reports it)
_SPILL_CACHE = {}       # device index -> uint8 tensor reused between calls (allocating 80 GB is not free)


def rho_spill_buffer(N: int, rank: int, world: int, nbins: int, dev, reserve_bytes: int = 0, max_tiles=None):
    """Buffer for the rho spill of one shard (include/pymn_b200.h, mn_corr_spill_layout).

    Pass 1 can leave the bin position of every cell pair in HBM (4 B per pair) so that pass 2 streams
    them back instead of recomputing the tiles on the tensor cores.  Returns (buffer or None,
    spill_tiles): as many tiles as fit into the free device memory minus ``reserve_bytes`` and a
    safety margin; the rest is recomputed.  ``MN_SPILL_GB`` caps the buffer (0 disables the spill)."""
    n_tiles, tile_bytes = _lib.corr_spill_layout(N, 0, N, rank, world, nbins)
    return _spill_alloc(n_tiles, tile_bytes, dev, reserve_bytes, max_tiles)


def _spill_alloc(n_tiles: int, tile_bytes: int, dev, reserve_bytes: int = 0, max_tiles=None):
    """(cached uint8 buffer or None, granted tiles) for a spill of up to n_tiles tiles."""
    total_tiles = n_tiles
    if max_tiles is not None:
        n_tiles = min(n_tiles, int(max_tiles))
    cap = _os.environ.get("MN_SPILL_GB")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    cached = _SPILL_CACHE.get(idx)
    have = int(cached.numel()) if cached is not None else 0
    free, _total = torch.cuda.mem_get_info(idx)
    # memory torch has cached but not handed out is reusable too
    free += torch.cuda.memory_reserved(idx) - torch.cuda.memory_allocated(idx)
    budget = free + have - int(reserve_bytes) - (4 << 30)
    if cap is not None:
        budget = min(budget, int(float(cap) * (1 << 30)))
    want_tiles = max(0, min(n_tiles, budget // tile_bytes))
    LAST_SPILL.update(total_tiles=total_tiles, tile_bytes=tile_bytes,
                      spill_tiles=int(want_tiles) if want_tiles >= (8 if max_tiles is None else 1) else 0)
    if want_tiles < (8 if max_tiles is None else 1):     # too small to pay for a second kernel
        return None, 0
    need = want_tiles * tile_bytes
    if have < need:
        _SPILL_CACHE.pop(idx, None)
        del cached
        torch.cuda.empty_cache()
        cached = torch.empty((need,), dtype=torch.uint8, device=dev)
        _SPILL_CACHE[idx] = cached
    return cached, want_tiles


def release_spill_cache():
    _SPILL_CACHE.clear()
    torch.cuda.empty_cache()


def corr_network_votes(rc: RankedCells, label: torch.Tensor, L: int, nbins=DEFAULT_NBINS, shard=(0, 1),
                       reduce_fn=None, spill_tiles=None, spill=None):
    """K4/K5: two tensor-core passes over the (never stored) cell x cell network.
    ``shard`` = (rank, world): this GPU handles every world-th 128-row tile; ``reduce_fn``
    (an in-place sum all-reduce over ranks) is applied to the histogram before the LUT and
    to votes / degree afterwards -- integer sums, so the result is bit-identical for any
    world size.  ``spill_tiles``: None = spill as many tiles as fit (rho_spill_buffer), an int caps the
    number of spilled tiles (0 = recompute everything; the result does not depend on it).
    Returns (votes uint64-as-int64 [N, L], degree [N], hist [nbins], lut)."""
    dev = rc.d_hi.device
    N = rc.M
    rank, world = shard
    zbuf = torch.zeros((nbins + N * L + N,), dtype=torch.int64, device=dev)       # one fill for the three accumulators
    hist, votes, degree = zbuf[:nbins], zbuf[nbins:nbins + N * L].view(N, L), zbuf[nbins + N * L:]
    if spill is not None:
        spill, n_sp = spill
    else:
        spill, n_sp = rho_spill_buffer(N, rank, world, nbins, dev, reserve_bytes=N * (L + 1) * 8, max_tiles=spill_tiles)
    _lib.call("mn_corr_hist", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, 0, N, rank, world, rc.G, nbins, hist,
              spill, n_sp)
    if reduce_fn is not None:
        reduce_fn(hist)
    lut = torch.empty((nbins + 2,), dtype=torch.int32, device=dev)
    _lib.call("mn_corr_lut", hist, nbins, N, lut)
    _lib.call("mn_corr_vote", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, 0, N, rank, world, rc.G, nbins, lut,
              label, L, votes, degree, spill, n_sp)
    if reduce_fn is not None:
        reduce_fn(votes)
        reduce_fn(degree)
    return votes, degree, hist, lut


def dist_shard():
    """(rank, world, reduce_fn) for the row-tile sharding of the full-network path.
    With torch.distributed initialised on >1 ranks (one process per GPU, NCCL) the two
    tensor-core passes are split by row tile and the integer partials summed with
    all-reduce; otherwise a single GPU does everything."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            def _sum(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return dist.get_rank(), dist.get_world_size(), _sum
    except Exception:
        pass
    return 0, 1, None


class StageTimer:
    """CUDA-event timer per named stage on the current stream (bench.py's roofline numbers)."""

    def __init__(self):
        self.events = {}

    def start(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.setdefault(name, []).append([ev, None])

    def stop(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events[name][-1][1] = ev

    def mean_ms(self):
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self.events.items()}


def us_default_device(Xd, perm, cell_label, seg_off, L, S, ndn=True, nbins=DEFAULT_NBINS, shard=None, timer=None,
                      spill_tiles=None):
    """Device-resident full-network pipeline: every argument already lives in HBM.
    Returns (auroc float64 [L, L] rows = test labels, n_pos, flags, hist)."""
    t = timer
    if t: t.start("rank_normalize")
    rc = rank_normalize(Xd, row_index=perm, drop_mode=0)
    if t: t.stop("rank_normalize")
    return network_from_ranked(rc, cell_label, seg_off, L, S, ndn, nbins, shard, timer, spill_tiles)


def network_from_ranked(rc: RankedCells, cell_label, seg_off, L, S, ndn=True, nbins=DEFAULT_NBINS, shard=None,
                        timer=None, spill_tiles=None):
    """Pass 1 -> LUT -> pass 2 -> votes -> AUROC on ranked cells that are already in label-sorted order."""
    rank, world, reduce_fn = shard if shard is not None else dist_shard()
    t = timer
    dev = rc.d_hi.device
    N = rc.M
    hist = torch.zeros((nbins,), dtype=torch.int64, device=dev)
    spill, n_sp = rho_spill_buffer(N, rank, world, nbins, dev, reserve_bytes=N * (L + 1) * 8 + N * L * 4 + (1 << 30),
                                   max_tiles=spill_tiles)
    if t: t.start("corr_hist")
    _lib.call("mn_corr_hist", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, 0, N, rank, world, rc.G, nbins, hist,
              spill, n_sp)
    if t: t.stop("corr_hist")
    if reduce_fn is not None:
        reduce_fn(hist)
    lut = torch.empty((nbins + 2,), dtype=torch.int32, device=dev)
    _lib.call("mn_corr_lut", hist, nbins, N, lut)
    votes = torch.zeros((N, L), dtype=torch.int64, device=dev)
    degree = torch.zeros((N,), dtype=torch.int64, device=dev)
    if t: t.start("corr_vote")
    _lib.call("mn_corr_vote", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, 0, N, rank, world, rc.G, nbins, lut,
              cell_label, L, votes, degree, spill, n_sp)
    if t: t.stop("corr_vote")
    if reduce_fn is not None:
        reduce_fn(votes)
        degree = None                     # = row sums of the reduced votes (mn_votes_finalize): no second collective
    vmat = torch.empty((L, N), dtype=torch.float32, device=dev)
    _lib.call("mn_votes_finalize", votes, degree, cell_label, N, L, int(bool(ndn)), vmat, N)
    if t: t.start("auroc")
    if reduce_fn is not None and world > 1 and L >= world:
        # every rank holds the reduced votes: the train columns (independent sorts) are dealt round-robin and
        # the column slices of the table all-gathered, as in the centroid path
        cols = torch.arange(rank, L, world, device=dev)
        auc_loc, n_pos = auroc(vmat.index_select(0, cols).contiguous(), int(cols.numel()), None, None, seg_off, S,
                               cell_label, L)
        auc = _gather_columns(auc_loc, L, rank, world)
    else:
        auc, n_pos = auroc(vmat, L, None, None, seg_off, S, cell_label, L)
    if t: t.stop("auroc")
    return auc, n_pos, rc.flags, hist


import time as _time
TRACE_UPLOAD = bool(_os.environ.get("MN_TRACE_UPLOAD"))       # diagnostic: timing-enabled chunk events, logged
UPLOAD_LOG = []


class AsyncUpload:
    """A dense host matrix on its way to the device: row chunks copied LAST ROWS FIRST on a side stream,
    one event per chunk.  ``chunks`` = [(row_lo, row_hi, event)] in upload order."""

    def __init__(self, Xd, host, side, pending):
        self.Xd, self._host, self._side, self._pending = Xd, host, side, list(pending)
        self.chunks = []
        self.shape = tuple(Xd.shape)

    def enqueue(self, n_chunks=None):
        """Queue the next ``n_chunks`` row chunks (all that are left if None) on the side stream.
        The copy engine serves transfers in issue order, whatever their stream: everything queued
        here delays the small label arrays the first kernels wait for, so start_upload() queues only
        the part that the host's label encoding hides and the pipeline queues the rest afterwards."""
        n = len(self._pending) if n_chunks is None else min(int(n_chunks), len(self._pending))
        with torch.cuda.stream(self._side):
            for _ in range(n):
                lo, hi = self._pending.pop(0)
                self.Xd[lo:hi].copy_(self._host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event(enable_timing=TRACE_UPLOAD)
                ev.record(self._side)
                self.chunks.append((lo, hi, ev))
                if TRACE_UPLOAD:
                    UPLOAD_LOG.append((lo, hi, ev, _time.perf_counter()))

    def upload_small(self, arrays):
        """int32 device copies of small host arrays, copied on the upload stream (FIFO with the matrix chunks queued
        so far); the current stream waits for them."""
        cur = torch.cuda.current_stream()
        dev = self.Xd.device
        hosts = [torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int32))).pin_memory() for a in arrays]
        outs = [torch.empty(h.shape, dtype=torch.int32, device=dev) for h in hosts]      # allocated for the current stream
        self._side.wait_stream(cur)              # (their blocks may have been in use by work queued on the current stream)
        with torch.cuda.stream(self._side):
            for o, h in zip(outs, hosts):
                o.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self._side)
        cur.wait_event(ev)
        self._small = hosts                      # keep the pinned sources alive until the call is over
        return outs

    def event_for_rows_from(self, row_lo: int):
        """Event after which every input row >= row_lo is on the device."""
        self.enqueue()
        ev = self.chunks[-1][2]
        for lo, _hi, e in self.chunks:
            if lo <= row_lo:
                ev = e
                break
        return ev

    def wait_all(self):
        self.enqueue()
        torch.cuda.current_stream().wait_event(self.chunks[-1][2])
        return self.Xd


class StagedUpload(AsyncUpload):
    """The same for a PAGEABLE host matrix: a worker thread copies the rows (last rows first) through the two pinned
    64 MB staging buffers and queues each piece on the upload stream -- the host copy of piece i + 1 runs while the
    DMA engine drains piece i, and both run while the main thread encodes the labels and launches the kernels of the
    chunks that have arrived.  (torch's large host copies release the GIL.)"""

    def __init__(self, Xd, host, side, pending):
        import threading
        super().__init__(Xd, host, side, pending)
        self._cond = threading.Condition()
        self._err = None
        self._done = False
        self._dev = Xd.device
        self._thread = threading.Thread(target=self._run, name="pymn_b200-upload", daemon=True)
        self._thread.start()

    def _run(self):
        try:
            torch.cuda.set_device(self._dev)
            idx = self._dev.index if self._dev.index is not None else torch.cuda.current_device()
            if idx not in _STAGE:
                _STAGE[idx] = (torch.empty((2, STAGE_BYTES), dtype=torch.uint8, pin_memory=True),
                               [torch.cuda.Event(), torch.cuda.Event()])
            stage, evs = _STAGE[idx]
            row_bytes = int(self._host.shape[1]) * 4
            rows_per = max(1, STAGE_BYTES // row_bytes)
            src = self._host.view(torch.uint8).view(self._host.shape[0], row_bytes)
            dst = self.Xd.view(torch.uint8).view(self.Xd.shape[0], row_bytes)
            i = 0
            for lo, hi in list(self._pending):
                for r0 in range(hi, lo, -rows_per):                  # last rows first inside the chunk too
                    r1, r0 = r0, max(lo, r0 - rows_per)
                    b = i & 1
                    i += 1
                    evs[b].synchronize()                            # the DMA that last used this buffer has drained
                    m = (r1 - r0) * row_bytes
                    stage[b, :m].view(r1 - r0, row_bytes).copy_(src[r0:r1])
                    with torch.cuda.stream(self._side):
                        dst[r0:r1].copy_(stage[b, :m].view(r1 - r0, row_bytes), non_blocking=True)
                        evs[b].record(self._side)
                ev = torch.cuda.Event(enable_timing=TRACE_UPLOAD)
                ev.record(self._side)
                with self._cond:
                    self.chunks.append((lo, hi, ev))
                    if TRACE_UPLOAD:
                        UPLOAD_LOG.append((lo, hi, ev, _time.perf_counter()))
                    self._cond.notify_all()
        except BaseException as e:                                  # noqa: BLE001 -- re-raised on the caller's thread
            self._err = e
        finally:
            with self._cond:
                self._done = True
                self._pending = []
                self._cond.notify_all()

    def enqueue(self, n_chunks=None):
        return                                                      # the worker queues everything

    def _check(self):
        if self._err is not None:
            raise self._err

    def event_for_rows_from(self, row_lo: int):
        with self._cond:
            while True:
                self._check()
                for lo, _hi, e in self.chunks:
                    if lo <= row_lo:
                        return e
                if self._done:
                    self._check()
                    return self.chunks[-1][2]
                self._cond.wait(0.05)

    def wait_all(self):
        self._thread.join()
        self._check()
        torch.cuda.current_stream().wait_event(self.chunks[-1][2])
        return self.Xd


# Measured at config 2 (1.6 GB pinned matrix, B200, ~55 GB/s H2D), median of 8 calls of the public API:
# upload-then-compute 123.5 ms; 16 upload chunks x 4 row chunks of the passes 109.8 ms (x 8: 112.1, x 16: 114.7).
# The first streamed version gained nothing: all chunk copies were queued up front and the copy engine, which
# serves transfers in issue order across streams, made the small label arrays -- which the first kernel
# needs -- wait for the whole matrix.  MN_STREAM_CHUNKS=1 turns streaming off.
SHARD_UPLOAD = _os.environ.get("MN_SHARD_UPLOAD", "1") != "0"     # multi-GPU: each rank uploads only its slice of the cells
STAGED_UPLOAD = _os.environ.get("MN_STAGED_UPLOAD", "1") != "0"   # pageable matrices: worker-thread staging, pipelined like pinned ones
STREAM_MIN_BYTES = 64 << 20
STREAM_EARLY_BYTES = 400 << 20          # queued before the host encodes labels (~8 ms of PCIe Gen5)
STREAM_CHUNKS = int(_os.environ.get("MN_STREAM_CHUNKS", "16"))
# row chunks of the passes (0 = as many as uploads).  Every chunk costs ~0.6 ms of launch tails, but the LAST one -- the
# first rows against all columns, the largest -- can only start when the whole matrix has arrived: on boxes whose
# H2D link runs at ~27 GB/s instead of ~55 that wait dominates.  Modelled / measured end-to-end times at config 2:
# 55 GB/s: 4 chunks 108.0 ms, 6: 107.6, 8: 108.8;  27 GB/s: 124.1 / 120.3 / 118.8.  6 is neutral on the fast link.
STREAM_DEV_CHUNKS = int(_os.environ.get("MN_STREAM_DEV_CHUNKS", "6"))


_XD_CACHE = {}          # device index -> uint8 buffer holding the streamed upload of the current call
_SIDE_STREAMS = {}      # device index -> the upload stream


def start_upload(X, full_network: bool = False):
    """Begin moving the expression matrix to the device.  A large dense float32 matrix in PINNED host memory
    is sent in row chunks, last rows first, on a side stream (AsyncUpload): the full-network pipeline
    starts ranking and correlating the last cells while the first ones are still crossing PCIe
    (us_default_streamed).  Anything else goes through to_device_matrix."""
    rank, world, _ = dist_shard()
    if (world > 1 and SHARD_UPLOAD and isinstance(X, np.ndarray) and X.dtype == np.float32 and X.ndim == 2
            and X.flags.c_contiguous and X.shape[0] >= 4 * world):
        lo, hi, per = shard_rows(X.shape[0], rank, world)
        sh = ShardedUpload(to_device(X[lo:hi], torch.float32), lo, hi, tuple(X.shape), rank, world, per)
        if full_network:
            # K1 is row-wise and the plane exchange needs no labels: both run (input order) while the host
            # still encodes the labels; ranked_cells_sharded only applies the label permutation afterwards
            sh.gathered = _rank_and_gather(sh, 0)
        return sh
    if (STREAM_CHUNKS > 1 and isinstance(X, np.ndarray) and X.dtype == np.float32 and X.ndim == 2
            and X.flags.c_contiguous and X.nbytes >= STREAM_MIN_BYTES):
        t = torch.from_numpy(X)
        if t.is_pinned() or STAGED_UPLOAD:
            staged = not t.is_pinned()
            dev = _dev()
            N = int(X.shape[0])
            # The device copy lives in a buffer kept between calls.  A fresh tensor per call would have to be marked
            # as used by the side stream (record_stream), which delays the reuse of its block: every other call then
            # pays a cudaMalloc of the whole matrix (observed: calls alternating between 122 and 162 ms).  Ordering is
            # by streams instead: the side stream waits for everything queued on the current one (the previous
            # call's kernels, which read the buffer) before the first chunk is copied.
            nbytes = int(X.nbytes)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            buf = _XD_CACHE.get(idx)
            if buf is None or buf.numel() < nbytes:
                _XD_CACHE.pop(idx, None)
                buf = None
                buf = _XD_CACHE[idx] = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            Xd = buf[:nbytes].view(torch.float32).view(tuple(X.shape))
            side = _SIDE_STREAMS.get(idx)
            if side is None:
                side = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            bounds = [int(round(k * N / STREAM_CHUNKS)) for k in range(STREAM_CHUNKS + 1)]
            pending = [(bounds[k], bounds[k + 1]) for k in reversed(range(STREAM_CHUNKS)) if bounds[k + 1] > bounds[k]]
            if staged:
                return StagedUpload(Xd, t, side, pending)
            up = AsyncUpload(Xd, t, side, pending)
            up.enqueue(max(1, int(round(STREAM_EARLY_BYTES / max(1, X.nbytes // STREAM_CHUNKS)))))
            return up
    return to_device_matrix(X)


def us_default_streamed(up: AsyncUpload, lay: LabelLayout, L, S, ndn=True, nbins=DEFAULT_NBINS, shard=None):
    """The full-network pipeline overlapped with the upload.  The network passes cover the pairs i < j, so
    the cells of device rows [b, b') only need rows >= b: processing the row chunks from the LAST to the
    first, chunk k can be ranked (K1) and correlated (pass 1) as soon as the input rows of device rows
    >= b_k have arrived -- with the usual study-sorted AnnData that is while earlier studies are still
    in flight (an arbitrary cell order degenerates to upload-then-compute, never to a wrong result).
    Integer histogram / vote sums make the result identical to us_default_device."""
    rank, world, reduce_fn = shard if shard is not None else dist_shard()
    dev = up.Xd.device
    N, G = up.shape
    cur = torch.cuda.current_stream()
    perm_np = np.asarray(lay.perm)
    # The label arrays travel on the UPLOAD stream, between the chunks queued so far and the rest of the matrix: the
    # copy engine drains one stream's queue before it serves another's, so a copy issued on the compute stream while
    # matrix chunks are in flight would only run after the whole 1.6 GB (the first kernel then starts 20 ms late).
    perm, cell_label, seg_off = up.upload_small([lay.perm, lay.cell_label, lay.seg_off])
    up.enqueue()                                 # the rest of the matrix follows the small arrays through the copy engine
    nd = max(1, min(STREAM_DEV_CHUNKS if STREAM_DEV_CHUNKS > 0 else STREAM_CHUNKS, N // 4096))
    # (row chunks start on multiples of 512: the tile height of the widest GEMM form)
    bounds = sorted(set([0, N] + [min(N, int(round(k * N / nd / 512.0)) * 512) for k in range(1, nd)]))
    nd = len(bounds) - 1
    # lowest input row that device rows >= b need
    suf_min = np.minimum.accumulate(perm_np[::-1])[::-1]
    rc = alloc_ranked(N, G, dev)
    layouts = [_lib.corr_spill_layout(N, bounds[k], bounds[k + 1], rank, world, nbins) for k in range(nd)]
    tile_bytes = layouts[0][1]
    spill, granted = _spill_alloc(sum(t for t, _ in layouts), tile_bytes, dev,
                                  reserve_bytes=N * (L + 1) * 8 + N * L * 4 + (1 << 30))
    hist = torch.zeros((nbins,), dtype=torch.int64, device=dev)
    plan = []                                    # (row0, row1, spill slice or None, tiles) in processing order
    off = 0
    for k in reversed(range(nd)):
        t_k = min(layouts[k][0], max(0, granted - off))
        sl = spill[off * tile_bytes:] if (spill is not None and t_k > 0) else None
        plan.append((bounds[k], bounds[k + 1], sl, t_k))
        off += t_k
    for r0, r1, sl, t_k in plan:
        cur.wait_event(up.event_for_rows_from(int(suf_min[r0])))
        rank_normalize(up.Xd, row_index=perm[r0:r1], drop_mode=0, out=rc, out_row0=r0)
        _lib.call("mn_corr_hist", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, r0, r1, rank, world, G, nbins, hist,
                  sl, t_k)
    if reduce_fn is not None:
        reduce_fn(hist)
    lut = torch.empty((nbins + 2,), dtype=torch.int32, device=dev)
    _lib.call("mn_corr_lut", hist, nbins, N, lut)
    votes = torch.zeros((N, L), dtype=torch.int64, device=dev)
    degree = torch.zeros((N,), dtype=torch.int64, device=dev)
    for r0, r1, sl, t_k in plan:
        _lib.call("mn_corr_vote", rc.d_hi, rc.d_lo, rc.ldd, rc.inv_norm, N, r0, r1, rank, world, G, nbins, lut,
                  cell_label, L, votes, degree, sl, t_k)
    if reduce_fn is not None:
        reduce_fn(votes)
        degree = None                     # = row sums of the reduced votes (mn_votes_finalize): no second collective
    vmat = torch.empty((L, N), dtype=torch.float32, device=dev)
    _lib.call("mn_votes_finalize", votes, degree, cell_label, N, L, int(bool(ndn)), vmat, N)
    auc, n_pos = auroc(vmat, L, None, None, seg_off, S, cell_label, L)
    return auc, n_pos, rc.flags, hist


def shard_rows(N: int, rank: int, world: int):
    """Rows [lo, hi) of the input a rank ingests, and the common block size `per` (the all-gather needs equal
    blocks: rank r's block starts at r * per, so gathered row j is input row j; the tail is padding)."""
    per = (int(N) + world - 1) // world
    return min(int(N), rank * per), min(int(N), (rank + 1) * per), per


def gather_rows(local: torch.Tensor, per: int, world: int) -> torch.Tensor:
    """All-gather per-rank row blocks (zero-padded to `per` rows) into [world * per, ...]."""
    import torch.distributed as dist
    buf = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return out


class ShardedUpload:
    """Multi-GPU ingestion: this rank uploads only ITS slice of the cells (rows [lo, hi) of the host matrix,
    input order); the ranked fp16 planes are exchanged over NVLink afterwards (us_default_sharded)."""

    def __init__(self, X_local, lo, hi, shape, rank, world, per):
        self.X_local, self.lo, self.hi, self.shape = X_local, lo, hi, shape
        self.rank, self.world, self.per = rank, world, per
        self.gathered = None          # (drop_mode, planes and scales of ALL cells in input order), see start_upload


def _rank_and_gather(sh: "ShardedUpload", drop_mode: int):
    """K1 on the own slice (input order) + all-gather of planes, scales and flags: (drop_mode, d_hi, d_lo, inv_norm,
    flags, ldd) with row j = input cell j."""
    rc_loc = rank_normalize(sh.X_local, drop_mode=drop_mode)
    return (drop_mode, gather_rows(rc_loc.d_hi, sh.per, sh.world),
            gather_rows(rc_loc.d_lo, sh.per, sh.world) if rc_loc.d_lo is not None else None,
            gather_rows(rc_loc.inv_norm, sh.per, sh.world), gather_rows(rc_loc.flags, sh.per, sh.world), rc_loc.ldd)


def ranked_cells_sharded(sh: ShardedUpload, lay: LabelLayout, drop_mode=0) -> RankedCells:
    """Data-parallel ingestion: every rank ranks the cells of its own input slice (K1, input
    order), one NCCL all-gather makes the fp16 planes (0.8 GB at config 2, ~ms over NVLink instead of N x
    the matrix over PCIe) and scales available everywhere, and a row gather puts them into label-sorted
    order.  K1 is row-wise, so the planes are identical to the single-GPU ones."""
    N, G = sh.shape
    g = sh.gathered if (sh.gathered is not None and sh.gathered[0] == drop_mode) else _rank_and_gather(sh, drop_mode)
    sh.gathered = None
    _, d_hi, d_lo, inv_norm, flags, ldd = g
    perm = to_device(np.asarray(lay.perm, dtype=np.int64))
    return RankedCells(d_hi.index_select(0, perm), d_lo.index_select(0, perm) if d_lo is not None else None, ldd,
                       inv_norm.index_select(0, perm), flags.index_select(0, perm), N, G)


def us_default_sharded(sh: ShardedUpload, lay: LabelLayout, L, S, ndn=True, nbins=DEFAULT_NBINS):
    """Full-network pipeline with data-parallel ingestion: ranked_cells_sharded, then the passes sharded by
    row tile as before; the tables are identical to the single-GPU ones."""
    rc = ranked_cells_sharded(sh, lay, 0)
    return network_from_ranked(rc, _i32(lay.cell_label), _i32(lay.seg_off), L, S, ndn, nbins, dist_shard())


def us_default(X, lay: LabelLayout, ndn=True, nbins=DEFAULT_NBINS):
    """metaNeighborUS_default.  Returns dict(auroc [L, L] rows = TEST labels, cols = TRAIN
    labels, both in row order -- the caller transposes / reorders, MetaNeighborUS.py:213).
    X: host / device matrix, CsrMatrix, or an AsyncUpload from start_upload()."""
    L, S = len(lay.row_labels), len(lay.studies)
    if isinstance(X, ShardedUpload):
        auc, n_pos, flags, hist = us_default_sharded(X, lay, L, S, ndn, nbins)
    elif isinstance(X, AsyncUpload):
        auc, n_pos, flags, hist = us_default_streamed(X, lay, L, S, ndn, nbins)
    else:
        Xd = to_device_matrix(X)
        auc, n_pos, flags, hist = us_default_device(Xd, _i32(lay.perm), _i32(lay.cell_label), _i32(lay.seg_off),
                                                    L, S, ndn, nbins)
    return dict(auroc=auc.cpu().numpy(), n_pos=n_pos.cpu().numpy(), flags=flags.cpu().numpy(),
                hist_total=int(hist.sum().item()))


# ---- supervised ------------------------------------------------------------
@dataclass
class SupervisedContext:
    """Everything the per-gene-set pipeline needs that does not depend on the gene set, resident in HBM:
    the expression matrix and the label arrays (uploading them again for every set costs four
    synchronous pageable copies per set)."""
    Xd: torch.Tensor | None
    perm: torch.Tensor
    cell_label: torch.Tensor
    cell_study: torch.Tensor
    lab_row_off: torch.Tensor
    seg_off: torch.Tensor
    n_studies: int
    max_label_rows: int = 0
    csc: CscMatrix | None = None       # sparse input: gene-set columns are expanded per set (csc_columns_dense)


def supervised_prepare(X, lay: LabelLayout) -> SupervisedContext:
    """Upload the shared expression matrix and the label layout once (MetaNeighbor.py:80-83)."""
    from scipy import sparse
    counts = np.diff(np.asarray(lay.lab_row_off))
    if sparse.issparse(X):
        return SupervisedContext(None, _i32(lay.perm), _i32(lay.cell_label), _i32(lay.cell_study),
                                 _i32(lay.lab_row_off), _i32(lay.seg_off), len(lay.studies),
                                 int(counts.max()) if len(counts) else 0, to_device_csc(X))
    return SupervisedContext(to_device(X, torch.float32), _i32(lay.perm), _i32(lay.cell_label), _i32(lay.cell_study),
                             _i32(lay.lab_row_off), _i32(lay.seg_off), len(lay.studies),
                             int(counts.max()) if len(counts) else 0)


def supervised_geneset(ctx: SupervisedContext, col_index: torch.Tensor, n_types: int, ndn=True, fast=False,
                       nbins=DEFAULT_NBINS, poison_out: torch.Tensor = None, out: torch.Tensor = None, spill=None):
    """One gene set.  ``col_index``: int32 device tensor of the set's gene columns.  The layout behind ``ctx``
    is that of JOINT labels study*K + type (all S*K ids, empty ranges allowed).  Returns the float64
    [S*K, K] AUROC table (rows = joint test labels) ON THE DEVICE (written into ``out`` when given): nothing here
    synchronises, so the caller can keep enqueuing gene sets and fetch all tables with one copy.

    ``poison_out`` (uint8 [1] tensor, fast path): set when some cell is constant and NON-ZERO over the gene
    set.  score_low_mem only drops cells whose sum is <= 0 (MetaNeighbor.py:121), so such a cell normalises
    to a NaN row (utils.py:146-147) that poisons `voters @ voter_id` and the candidates' ranks: the
    reference's result for the gene set is NaN for every cell type.  K1 drops the cell
    (flag bit 0 without bit 1); the caller reproduces the NaN column from this flag.
    ``spill``: (buffer, tiles) of the rho spill decided once per run (None = decide here)."""
    S = ctx.n_studies
    K = n_types
    if ctx.csc is not None:
        # sparse input: the set's columns as a dense [N, g] tile (the matrix itself never densifies)
        rc = rank_normalize(csc_columns_dense(ctx.csc, col_index), row_index=ctx.perm, drop_mode=1 if fast else 0,
                            limbs=bool(fast))
    else:
        rc = rank_normalize(ctx.Xd, row_index=ctx.perm, col_index=col_index, drop_mode=1 if fast else 0,
                            limbs=bool(fast))
    dev = rc.device
    cell_label, cell_study = ctx.cell_label, ctx.cell_study
    vmat = torch.empty((K, rc.M), dtype=torch.float32, device=dev)
    if fast:
        C, n_valid = centroids(rc, ctx.lab_row_off, S * K, max_label_rows=ctx.max_label_rows)
        votes = votes_exact(rc, C, None, S * K)
        _lib.call("mn_loso_finalize", votes, rc.M, 0, None, n_valid, cell_study, rc.M, S, K, int(bool(ndn)), vmat, rc.M)
        test_label = mask_dropped(cell_label, rc.inv_norm, rc.flags, poison_out)
    else:
        votes, degree, _, _ = corr_network_votes(rc, cell_label, S * K, nbins, spill=spill)
        _lib.call("mn_loso_finalize", votes, S * K, 1, degree, None, cell_study, rc.M, S, K, int(bool(ndn)), vmat, rc.M)
        test_label = cell_label
    auc, _ = auroc(vmat, K, None, None, ctx.seg_off, S, test_label, S * K, out=out)
    return auc


SUP_STREAMS = int(_os.environ.get("MN_SUP_STREAMS", "2"))


def supervised_tables(ctx: SupervisedContext, all_cols: torch.Tensor, offs, n_types: int, ndn=True, fast=False,
                      nbins=DEFAULT_NBINS):
    """All gene sets of one rank, back to back on the current stream (MetaNeighbor.py:84-93).  ``all_cols``:
    int32 device tensor, the gene columns of set j are all_cols[offs[j]:offs[j + 1]].  Returns
    (tables float64 [n_sets, S*K, K] on the device, poison uint8 [n_sets] -- see supervised_geneset);
    nothing synchronises.  The rho spill of the full-network mode is sized once for the whole run (its size
    depends on N only; asking the driver for the free memory costs more than a gene set's kernels)."""
    n_sets = len(offs) - 1
    dev = ctx.perm.device
    S, K = ctx.n_studies, n_types
    N = int(ctx.perm.numel())
    tables = torch.empty((n_sets, S * K, K), dtype=torch.float64, device=dev)
    poison = torch.zeros((max(1, n_sets),), dtype=torch.uint8, device=dev)
    if fast and n_sets > 0 and ctx.csc is None:
        # one C-ABI call for all sets: the host loop over the sets runs in C (mn_supervised_fast_sets)
        offs64 = np.ascontiguousarray(np.asarray(offs, dtype=np.int64))
        g_max = int(np.diff(offs64).max())
        wbytes = _lib.supervised_fast_workspace_bytes(N, g_max, S, K)
        ws = torch.empty((wbytes,), dtype=torch.uint8, device=dev)
        _lib.call("mn_supervised_fast_sets", ctx.Xd, int(ctx.Xd.stride(0)), ctx.perm, N, all_cols,
                  int(offs64.ctypes.data), n_sets, g_max, ctx.lab_row_off, ctx.cell_label, ctx.cell_study, ctx.seg_off,
                  S, K, int(bool(ndn)), int(ctx.max_label_rows), ws, wbytes, tables, poison)
        return tables, poison
    spill = None
    if not fast and n_sets > 0:
        spill = rho_spill_buffer(N, 0, 1, nbins, dev, reserve_bytes=N * (S * K + 1) * 8 + N * K * 4 + (2 << 30))
    n_streams = SUP_STREAMS if (not fast and n_sets >= 4 and spill is not None and spill[0] is not None) else 1
    if n_streams > 1:
        # Full-network mode, gene sets alternate between two streams (each with its own rho spill): the small
        # kernels of one set (K1, CDF table, LOSO combination, AUROC: ~0.13 of its 0.75 ms, none of which fills the
        # machine) run under the two network passes of the other.
        cur = torch.cuda.current_stream()
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        need = int(spill[1]) * int(LAST_SPILL["tile_bytes"])         # (the cached buffer may be far larger than this run needs)
        spills = [spill] + [(torch.empty((need,), dtype=torch.uint8, device=dev), spill[1]) for _ in range(n_streams - 1)]
        for st_ in streams:
            st_.wait_stream(cur)
        for j in range(n_sets):
            with torch.cuda.stream(streams[j % n_streams]):
                supervised_geneset(ctx, all_cols[int(offs[j]):int(offs[j + 1])], K, ndn, fast, nbins,
                                   out=tables[j], spill=spills[j % n_streams])
        for st_ in streams:
            cur.wait_stream(st_)
        return tables, poison
    for j in range(n_sets):
        supervised_geneset(ctx, all_cols[int(offs[j]):int(offs[j + 1])], K, ndn, fast, nbins,
                           poison_out=poison[j:j + 1] if fast else None, out=tables[j], spill=spill)
    return tables, poison
