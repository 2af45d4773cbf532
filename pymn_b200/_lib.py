"""ctypes binding of the C ABI in include/pymn_b200.h.

The shared library ``pymn_b200/libpymn_b200.so`` is built in-tree by
``pymn_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is no CPU
fallback: if the library is missing or a call fails, this module raises.
Tensors are torch CUDA tensors; only their ``data_ptr()`` and the current
CUDA stream cross the boundary (no torch types in the ABI).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYMN_B200_LIB") or os.path.join(_HERE, "libpymn_b200.so")   # (override: tuning builds)

_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32

# name -> (restype, argtypes); mirrors include/pymn_b200.h one to one
SIGNATURES = {
    "mn_last_error": (C.c_char_p, []),
    "mn_abi_version": (C.c_int, []),
    "mn_launch_count": (_i64, []),
    "mn_rank_normalize": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "mn_rank_normalize_csr": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "mn_rank_normalize_limbs": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "mn_rank_normalize_csr_limbs": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "mn_gene_stats": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "mn_csc_gather_dense": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _i64, _vp]),
    "mn_centroids": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "mn_centroids_fixed": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp]),
    "mn_centroids_fixed_i8": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp]),
    "mn_group_sum_fixed": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mn_centroids_finish": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "mn_group_sum": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "mn_split_f16": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp]),
    "mn_votes_fast": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "mn_limb_split_cells": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp]),
    "mn_limb_split_centroids": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i64, _i64, _vp, _vp, _vp]),
    "mn_votes_den64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "mn_votes_fast_i8": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _i64,
                                   _vp, _i64, _vp, _vp, _vp]),
    "mn_votes_fast_i8_peer": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32,
                                        _i64, _vp, _i64, _vp, _vp, _vp]),
    "mn_votes_den_i8": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "mn_corr_hist": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "mn_corr_spill_layout": (C.c_int, [_i64, _i64, _i64, _i32, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "mn_corr_lut": (C.c_int, [_vp, _i32, _i64, _vp, _vp]),
    "mn_corr_vote": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "mn_votes_finalize": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
    "mn_mask_dropped": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mn_loso_finalize": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp]),
    "mn_votes_unpack": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _vp, _i64, _vp]),
    "mn_supervised_fast_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "mn_supervised_fast_sets": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32,
                                          _i64, _vp, _i64, _vp, _vp, _vp]),
    "mn_auroc_workspace_bytes": (_i64, [_i32, _i64, _i32, _i32]),
    "mn_auroc": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _vp, _vp, _vp, _i64, _vp]),
    "mn_one_vs_best": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"pymn_b200: {LIB_PATH} is missing -- build it with pymn_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pymn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t):
    if t is None:
        return None
    # (a row-pitched matrix -- unit column stride, rows ld apart -- is fine: its pitch travels as an argument)
    assert t.is_cuda and (t.is_contiguous() or (t.dim() == 2 and t.stride(1) == 1)), "device, contiguous tensors only"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Call an int-returning entry point; tensors are converted to pointers,
    the current CUDA stream is appended; raises RuntimeError on failure."""
    lib = load()
    conv = [(_ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
    rc = getattr(lib, name)(*conv, _stream())
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {lib.mn_last_error().decode()}")


def launch_count() -> int:
    return int(load().mn_launch_count())


def corr_spill_layout(n: int, row0: int, row1: int, shard_rank: int, shard_world: int, nbins: int):
    """(number of tiles, bytes per tile) of one shard's rho spill."""
    lib = load()
    nt, tb = _i64(0), _i64(0)
    rc = lib.mn_corr_spill_layout(n, row0, row1, shard_rank, shard_world, nbins, C.byref(nt), C.byref(tb))
    if rc != 0:
        raise RuntimeError(f"mn_corr_spill_layout failed (rc={rc}): {lib.mn_last_error().decode()}")
    return int(nt.value), int(tb.value)


def supervised_fast_workspace_bytes(n: int, g_max: int, s: int, k: int) -> int:
    v = int(load().mn_supervised_fast_workspace_bytes(n, g_max, s, k))
    if v < 0:
        raise RuntimeError("mn_supervised_fast_workspace_bytes: bad shape")
    return v


def auroc_workspace_bytes(ncols: int, m: int, n_seg: int, n_labels: int) -> int:
    return int(load().mn_auroc_workspace_bytes(ncols, m, n_seg, n_labels))
