#!/usr/bin/env python
"""bench.py -- MetaNeighborUS cell-pairs/s (and MetaNeighbor gene-sets/s) on the BASELINE.json workloads.

    python bench.py --gpus N --steps K --warmup W                    # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Headline workload (BASELINE.json configs[1], the config the metric is quoted on and the largest full-network
case that fits one GPU): MetaNeighborUS fast_version=False, full Spearman cell x cell network, 4 studies x
50,000 cells x 2,000 HVGs, 40 clusters (N = 200,000; N^2 = 4e10 cell pairs), synthetic counts (pymn_b200/synth.py).

A "step" is one pass of the hot path over that input: per-cell rank + normalise -> pass 1 (rho histogram on
tensor cores + rho spill) -> CDF table -> pass 2 (rank-transformed votes + node degree) -> vote matrix ->
rank-sum AUROC table.  `value` = N^2 / step time with inputs resident in HBM.  `e2e` = the same metric through
the public API (pymn_b200.MetaNeighborUS on an AnnData-like object whose X is a pinned HOST array: label
encoding, H2D of X, all kernels, D2H of the table, DataFrame); `e2e_pageable` the same with an ordinary
(pageable) NumPy array, `e2e_cold` the very FIRST call of the process on that pageable array (library load,
80 GB spill allocation, staging buffers -- what a user who calls MetaNeighborUS once pays).

The one JSON line also carries
  parity        the GPU path against the reference's own table on the 4 x 1,000-cell sample of the same generator
                (tests/golden/big_default_c2_4k.npz, produced by the unmodified reference) and, at N = 1, against
                the oracle table the cpu_baseline leg computes in this very run;
  table_sha256  hash of the headline AUROC table: identical for 1 / 2 / 4 / 8 GPUs (integer partial sums);
  configs       the other BASELINE configs measured in the same run: c1 (fast), c3 (supervised, gene-sets/s,
                both modes), c4 (BICCN scale, one-vs-best), c5 (pretrained model, 1M query cells).

N > 1 (torchrun, one process per GPU): the two tensor-core passes are split by row tile across ranks, integer
partials are summed with NCCL (strong scaling: the workload is fixed); the centroid paths shard cells (K1) and
train columns (votes, AUROC); the supervised path shards gene sets.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "MetaNeighborUS cell-pairs/s (corr+vote+AUROC)"
UNIT = "cell-pairs/s"

CONFIGS = {
    # name: (studies, cells/study, genes, types/study, description)
    "c2": (4, 50000, 2000, 10, "c2: MetaNeighborUS fast_version=False, full network, 4 studies x 50,000 cells x 2,000 HVGs, 40 clusters"),
    "c2_small": (4, 5000, 2000, 10, "c2 at 1/10 cells (dev check only)"),
    "c2_mid": (4, 15000, 2000, 10, "c2 at 3/10 cells (profiling only)"),
}
# the other BASELINE configs, reported in the `configs` block: (studies, cells/study, genes, types/study, one_vs_best)
FAST_CONFIGS = {
    "c1": (2, 10000, 2000, 20, False, "c1: MetaNeighborUS fast_version=True, 2 studies x 10,000 cells x 2,000 HVGs x 20 cell types"),
    "c4": (7, 71429, 3000, 150, True, "c4: BICCN-scale MetaNeighborUS fast_version=True one_vs_best=True, 7 studies x 71,429 cells x 3,000 HVGs, 150 clusters per study (L = 1,050)"),
}
C5 = (1, 1000000, 3000, 50, "c5: trainModel (3 studies x 50 clusters x 50 cells = 150-column model) + MetaNeighborUS(trained_model=...) on 1,000,000 query cells x 3,000 genes, 50 query labels")
C3 = (3, 6667, 5000, 10, 1000, "c3: MetaNeighbor supervised, 1,000 random gene sets (10-200 genes) over 20,001 cells x 3 studies, 10 cell types, leave-one-study-out")
REF_SAMPLE_CELLS_PER_STUDY = 1000        # CPU arm: 4 x 1,000 cells of the same generator (N^2 = 1.6e7 pairs)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 100 ms via NVML while running.  NVML is initialised in
    the constructor, on the caller's thread and BEFORE any timed region: nvmlInit takes the driver lock for up to
    several hundred ms on some boxes, which stalls kernel launches if it happens inside the timed steps."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self._nv = self._h = self._get = None
        self._names = {}
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            self._get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)        # first queries are the slow ones
            self._get(self._h)
        except Exception as e:          # NVML unavailable: report what we have
            self._nv = None
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        nv = self._nv
        if nv is None:
            return
        try:
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = self._get(self._h)
                for bit, name in self._names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def mark(self):
        """The timed region starts: forget what was sampled during the warm-up."""
        self.samples = []
        self.reasons = {r for r in self.reasons if r.startswith("nvml_")}

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _labels(S, n, K):
    from pymn_b200.synth import make_labels
    s, c, _, _ = make_labels(S, n, K)
    return s, c


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on host cores
# ---------------------------------------------------------------------------
def cpu_reference_time(G, K, n_per_study, S=4, seed=2, repeats=1):
    """Times oracle.mn_oracle.metaneighbor_us_default (NumPy float64 restatement of
    MetaNeighborUS.py:145-224 + utils.py:35-75) on a bounded sample.  Returns (pairs/s, seconds, N, table)."""
    from oracle import mn_oracle as O
    from pymn_b200.synth import make_counts
    x, s, c = make_counts(seed, S, n_per_study, G, K)
    best, tab = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        tab = O.metaneighbor_us_default(x, s, c, True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n = x.shape[0]
    return n * n / best, best, n, tab


def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(n)) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, n, G, K, desc = CONFIGS[args.config]
    for _ in range(args.warmup):
        cpu_reference_time(G, K, REF_SAMPLE_CELLS_PER_STUDY, S)
    times = []
    pairs = None
    for _ in range(args.steps):
        v, dt, ncell, _ = cpu_reference_time(G, K, REF_SAMPLE_CELLS_PER_STUDY, S)
        times.append(dt)
        pairs = ncell * ncell
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms / 1e3)
    cores = _blas_threads()
    sample = (f"oracle port (NumPy float64) of the reference's full-network path on {S}x{REF_SAMPLE_CELLS_PER_STUDY} "
              f"cells x {G} genes of the same generator (N^2 = {pairs:.2e} pairs per step); the full workload "
              f"needs a 320 GB float64 N x N matrix and does not run on the host")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count(), "blas_threads": cores, "numpy": np.__version__},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs CUDA GPUs"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world} (launch N>1 with torchrun)"
        self.args = args
        self.peaks = _peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: int) -> int:
        t = self.torch.tensor([v], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def timed(self, fn, warmup, steps):
        """Device time of `steps` calls of fn(timer) after `warmup` untimed ones: CUDA events on the current
        stream, barrier + synchronize on both sides, max over ranks.  Returns (ms per step, stage ms, launches, last result)."""
        from pymn_b200 import _lib, engine
        import gc
        torch = self.torch
        for _ in range(warmup):
            fn(None)
        # as `timeit` does: no cyclic garbage collection inside the timed region (a full collection of this process's
        # heap -- a few million label strings -- takes 100-150 ms and used to land in the first timed step)
        gc.collect()
        gc_was = gc.isenabled()
        gc.disable()
        self.barrier()
        timer = engine.StageTimer()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        marks = [e0]
        for _ in range(steps):
            out = fn(timer)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        e1.record()
        self.barrier()
        if gc_was:
            gc.enable()
        launches = self.sum_over_ranks(_lib.launch_count() - l0)
        ms = self.max_over_ranks(e0.elapsed_time(e1)) / steps
        self.last_step_ms = [round(self.max_over_ranks(a.elapsed_time(b)), 3) for a, b in zip(marks[:-1], marks[1:])]
        return ms, timer.mean_ms(), launches, out

    def wall(self, fn, warmup, steps):
        """Host wall clock of the public API call (the result is a host table, so the call has synchronised)."""
        import gc
        for _ in range(warmup):
            fn()
        gc.collect()
        gc_was = gc.isenabled()
        gc.disable()
        self.barrier()
        t0 = time.perf_counter()
        out = None
        self.last_wall_ms = []
        for _ in range(steps):
            t1 = time.perf_counter()
            out = fn()
            self.last_wall_ms.append(round((time.perf_counter() - t1) * 1e3, 2))
        self.barrier()
        dt = time.perf_counter() - t0
        if gc_was:
            gc.enable()
        return self.max_over_ranks(dt / steps), out


def _adata(X, s, c, G):
    import pandas as pd
    from pymn_b200.anndata_lite import AnnDataLite
    obs = pd.DataFrame({"study_id": s, "cell_type": c})
    var = pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=pd.Index([f"g{i}" for i in range(G)], dtype=object))
    return AnnDataLite(X, obs, var)


def _hbm(ctx, nbytes, ms):
    ach = nbytes / (ms / 1e3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": ctx.peaks["hbm"], "unit": "GB/s", "frac": ach / ctx.peaks["hbm"],
            "ms_per_launch": ms, "traffic": None}


def _tensor(ctx, flop, ms):
    ach = flop / (ms / 1e3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": ctx.peaks["tf_sust"], "unit": "TFLOP/s",
            "frac": ach / ctx.peaks["tf_sust"], "ms_per_launch": ms, "traffic": None}


def parity_block(ctx, oracle_table=None):
    """The GPU path on the 4 x 1,000 x 2,000, L = 40 sample of the headline generator, against the table the
    UNMODIFIED reference produced for it (committed fixture) and, when the cpu_baseline leg ran, against the
    oracle table of this very run.  Per-cell tolerance max(1e-5, 2 / (n_pos n_neg))."""
    import pymn_b200
    from pymn_b200.synth import make_counts
    S, _, G, K, _ = CONFIGS["c2"]
    x, s, c = make_counts(2, S, REF_SAMPLE_CELLS_PER_STUDY, G, K)
    tab = pymn_b200.MetaNeighborUS(_adata(x, s, c, G), "study_id", "cell_type", fast_version=False,
                                   symmetric_output=False, save_uns=False)
    n_pos = REF_SAMPLE_CELLS_PER_STUDY // K
    quantum = 1.0 / (n_pos * (REF_SAMPLE_CELLS_PER_STUDY - n_pos))
    tol = max(1e-5, 2.0 * quantum) * (1.0 + 1e-9)        # (a difference of exactly two quanta is a rounded float64)
    out = {"sample": f"{S}x{REF_SAMPLE_CELLS_PER_STUDY} cells x {G} genes, L={S * K}, nbins=16384 (the CPU arm's sample)",
           "quantum": quantum, "tol": tol, "table_sha256": _sha(tab.values)}
    path = os.path.join(ROOT, "tests", "golden", "big_default_c2_4k.npz")
    ok = True
    if os.path.isfile(path):
        z = np.load(path, allow_pickle=False)
        same_labels = [str(v) for v in z["out_auroc_rows"]] == [str(v) for v in tab.index]
        d = float(np.nanmax(np.abs(tab.values - z["out_auroc_values"]))) if same_labels else float("inf")
        out.update(vs_reference_fixture={"max_abs": d, "ok": bool(same_labels and d <= tol)})
        ok = ok and same_labels and d <= tol
    if oracle_table is not None:
        d = float(np.nanmax(np.abs(tab.values - oracle_table[0])))
        same_labels = [str(v) for v in oracle_table[1]] == [str(v) for v in tab.index]
        out.update(vs_oracle_this_run={"max_abs": d, "ok": bool(same_labels and d <= tol)})
        ok = ok and same_labels and d <= tol
    out["max_abs"] = max([v["max_abs"] for k, v in out.items() if isinstance(v, dict)] or [float("nan")])
    out["ok"] = bool(ok)
    return out


def bench_fast_config(ctx, name):
    """c1 / c4: MetaNeighborUS fast_version=True, device-resident step + e2e through the API."""
    import pymn_b200
    from pymn_b200 import engine
    from pymn_b200.synth import make_counts_torch
    from pymn_b200.utils import label_layout
    torch = ctx.torch
    S, n, G, K, ovb, desc = FAST_CONFIGS[name]
    Xd, _, _ = make_counts_torch(4 if name == "c4" else 1, S, n, G, K, ctx.dev)
    s, c = _labels(S, n, K)
    lay = label_layout(s, c)
    N, L = int(Xd.shape[0]), len(lay.row_labels)
    ms, st, launches, out = ctx.timed(lambda t: engine.us_fast_device(Xd, lay, True, ovb, t), 2, 5)
    steps_ms = list(ctx.last_step_ms)
    table = out[0].cpu().numpy()
    Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
    Xh.copy_(Xd)
    torch.cuda.synchronize()
    del Xd
    ad = _adata(Xh.numpy(), s, c, G)
    e2e_s, tab = ctx.wall(lambda: pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=True, one_vs_best=ovb,
                                                           symmetric_output=False, save_uns=False), 1, 2)
    assert np.array_equal(tab.values, table, equal_nan=True), f"{name}: API table differs from the device-resident run"
    pairs = float(N) * float(N)
    bpe = 4.0 + (4.0 if G > 2048 else 2.0)
    kern = {"rank_normalize": _hbm(ctx, bpe * N * G, st["rank_normalize"]),
            "votes_gemm": _tensor(ctx, 2.0 * N * G * (L + S) / ctx.world, st["votes_gemm"]),
            "auroc": _hbm(ctx, (4.0 * N * L + 4.0 * N) / ctx.world, st["auroc"])}
    dom = max(("rank_normalize", "votes_gemm", "auroc"), key=lambda k: st[k])
    del Xh, ad
    return {"workload": desc, "metric": "MetaNeighborUS effective cell-pairs/s (N^2 / step; the pairs are never formed)",
            "unit": UNIT, "n_cells": N, "n_genes": G, "n_labels": L, "value": pairs / (ms / 1e3), "ms_per_step": ms,
            "steps_ms": steps_ms,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": N * G * 4 + 12 * N,
                    "d2h_bytes_per_step": L * L * 16 + N},
            "gpu_launches": launches, "stage_ms": st, "roofline": dict(kern[dom], kernel=dom), "kernels": kern,
            "table_sha256": _sha(table)}


def bench_c5(ctx):
    """c5: trainModel on a 150-cluster reference, then MetaNeighborUS(trained_model=...) on 1M query cells."""
    import pandas as pd
    import pymn_b200
    from pymn_b200 import engine
    from pymn_b200.synth import make_counts, make_counts_torch
    from pymn_b200.utils import label_layout
    torch = ctx.torch
    S, n, G, K, desc = C5
    xt, st_, ct_ = make_counts(6, 3, 2500, G, 50)          # reference atlas: 3 studies x 50 clusters x 50 cells
    t0 = time.perf_counter()
    model = pymn_b200.trainModel(_adata(xt, st_, ct_, G), "study_id", "cell_type")
    train_s = time.perf_counter() - t0
    Lt = model.shape[1]
    Xd, _, _ = make_counts_torch(5, S, n, G, K, ctx.dev)
    s, c = _labels(S, n, K)
    lay = label_layout(s, c, replace_bar=True)
    N, L = int(Xd.shape[0]), len(lay.row_labels)
    cent = engine.to_device(np.ascontiguousarray(model.iloc[1:].values.T, dtype=np.float32))
    n_cells = engine.to_device(np.asarray(model.iloc[0].values, dtype=np.float32))
    st_names = np.array([str(v).split("|")[0] for v in model.columns.values], dtype=object)
    st_uniq, st_codes = np.unique(st_names, return_inverse=True)
    mstudy = engine._i32(st_codes.astype(np.int32))
    ms, st, launches, out = ctx.timed(
        lambda t: engine.us_pretrained_device(Xd, lay, cent, n_cells, mstudy, len(st_uniq), True, False, t), 2, 5)
    steps_ms = list(ctx.last_step_ms)
    table = out[0].cpu().numpy()
    Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
    Xh.copy_(Xd)
    torch.cuda.synchronize()
    del Xd
    ad = _adata(Xh.numpy(), s, c, G)
    import warnings
    def call():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", trained_model=model, symmetric_output=False,
                                            save_uns=False)
    e2e_s, tab = ctx.wall(call, 1, 2)
    assert np.array_equal(tab.values, table, equal_nan=True), "c5: API table differs from the device-resident run"
    eff_pairs = float(N) * float(model.iloc[0].values.sum())
    kern = {"rank_normalize": _hbm(ctx, 8.0 * N * G, st["rank_normalize"]),
            "votes_gemm": _tensor(ctx, 2.0 * N * G * (Lt + len(st_uniq)) / ctx.world, st["votes_gemm"]),
            "auroc": _hbm(ctx, (4.0 * N * Lt + 4.0 * N) / ctx.world, st["auroc"])}
    dom = max(kern, key=lambda k: st[k])
    del Xh, ad
    return {"workload": desc, "metric": "MetaNeighborUS effective cell-pairs/s (N_query x sum of model n_cells / step)",
            "unit": UNIT, "n_query_cells": N, "n_genes": G, "model_columns": Lt, "value": eff_pairs / (ms / 1e3),
            "query_cells_per_s": N / (ms / 1e3), "ms_per_step": ms, "steps_ms": steps_ms, "train_model_ms": train_s * 1e3,
            "e2e": {"value": eff_pairs / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "query_cells_per_s": N / e2e_s, "h2d_bytes_per_step": N * G * 4 + 12 * N + Lt * G * 4,
                    "d2h_bytes_per_step": L * Lt * 16 + N},
            "gpu_launches": launches, "stage_ms": st, "roofline": dict(kern[dom], kernel=dom), "kernels": kern,
            "table_sha256": _sha(table)}


def bench_c3(ctx):
    """c3: MetaNeighbor supervised through the public API, default (full network per gene set) and fast_version."""
    import pandas as pd
    import pymn_b200
    from pymn_b200 import _lib
    from pymn_b200.anndata_lite import AnnDataLite
    from pymn_b200.synth import make_counts_torch, make_genesets
    torch = ctx.torch
    S, n, G, K, n_sets, desc = C3
    Xd, _, _ = make_counts_torch(3, S, n, G, K, ctx.dev)
    Xh = torch.empty(tuple(Xd.shape), dtype=torch.float32, pin_memory=True)
    Xh.copy_(Xd)
    torch.cuda.synchronize()
    del Xd
    s, c = _labels(S, n, K)
    genes = pd.Index([f"g{i}" for i in range(G)], dtype=object)
    ad = AnnDataLite(Xh.numpy(), pd.DataFrame({"study_id": s, "cell_type": c}), pd.DataFrame(index=genes))
    gmat = make_genesets(3, G, n_sets, 10, 200)
    gs = pd.DataFrame(gmat, index=genes, columns=[f"set{j}" for j in range(n_sets)])
    N = int(Xh.shape[0])
    out = {"workload": desc, "metric": "MetaNeighbor gene-sets/s", "unit": "gene-sets/s", "n_cells": N, "n_sets": n_sets,
           "sum_set_sizes": int(gmat.sum())}
    for mode, fast in (("default", False), ("fast_version", True)):
        pymn_b200.MetaNeighbor(ad, "study_id", "cell_type", gs.iloc[:, :8 * ctx.world], fast_version=fast, save_uns=False)
        ctx.barrier()
        l0 = _lib.launch_count()
        dt, tab = ctx.wall(lambda: pymn_b200.MetaNeighbor(ad, "study_id", "cell_type", gs, fast_version=fast, save_uns=False), 0, 1)
        launches = ctx.sum_over_ranks(_lib.launch_count() - l0)
        entry = {"value": n_sets / dt, "unit": "gene-sets/s", "seconds": dt, "gpu_launches": launches,
                 "e2e": {"value": n_sets / dt, "unit": "gene-sets/s", "h2d_bytes_per_step": N * G * 4 + int(gmat.sum()) * 4,
                         "d2h_bytes_per_step": n_sets * S * K * K * 8},
                 "mean_auroc": float(np.nanmean(tab.values)), "table_sha256": _sha(tab.values)}
        if not fast:
            # per gene set the path ranks all N^2 correlations: algorithmic contraction flops 2 N^2 g (upper triangle: half)
            entry["network_pairs_per_s"] = n_sets * float(N) * float(N) / dt
        out[mode] = entry
    out["value"] = out["default"]["value"]
    del Xh, ad
    return out


def run_ours(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    import pymn_b200
    from pymn_b200 import _lib, engine
    from pymn_b200.synth import make_counts_torch
    from pymn_b200.utils import label_layout
    world, rank, dev = ctx.world, ctx.rank, ctx.dev

    S, n, G, K, desc = CONFIGS[args.config]
    N = S * n
    pairs = float(N) * float(N)

    # synthetic inputs, identical on every rank (same seed)
    Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
    s_lab, c_lab = _labels(S, n, K)
    X_page = Xd.cpu().numpy()                      # an ordinary pageable NumPy array, as a user holds it
    adata_page = _adata(X_page, s_lab, c_lab, G)

    def api_call(ad):
        return pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=False, symmetric_output=False,
                                        save_uns=False)

    # ---- e2e_cold: the FIRST call of this process, pageable input (nothing of ours has touched the GPU yet)
    ctx.barrier()
    t0 = time.perf_counter()
    tab_cold = api_call(adata_page)
    cold_s = ctx.max_over_ranks(time.perf_counter() - t0)

    lay = label_layout(s_lab, c_lab)
    L = len(lay.row_labels)
    perm, cell_label, seg_off = engine._i32(lay.perm), engine._i32(lay.cell_label), engine._i32(lay.seg_off)
    shard = engine.dist_shard()

    # ---- device-resident timing ------------------------------------------------
    def step(timer):
        return engine.us_default_device(Xd, perm, cell_label, seg_off, L, S, True, engine.DEFAULT_NBINS, shard, timer)
    sampler = ClockSampler(ctx.local_rank)          # NVML initialised here and the thread started before the warm-up:
    sampler.start()                                 # nothing of it begins inside the timed region
    for _ in range(args.warmup):
        step(None)
    ctx.barrier()
    sampler.mark()
    ms_step, stage_ms, launches, out = ctx.timed(step, 0, args.steps)
    steps_ms = list(ctx.last_step_ms)
    clocks = sampler.stop()
    value = pairs / (ms_step / 1e3)
    auc_dev = out[0]
    assert bool(torch.isfinite(auc_dev).all()), "AUROC table has non-finite entries"
    order = np.argsort(lay.sorted_pos)
    ref_tab = auc_dev.cpu().numpy()[np.ix_(order, order)].T
    assert np.array_equal(tab_cold.values, ref_tab), "API table (cold call) differs from the device-resident run"
    sp = dict(engine.LAST_SPILL)

    # ---- end to end through the public API, host buffers -------------------------
    if os.environ.get("MN_BENCH_GC_FREEZE"):        # diagnostic only: take the bench's own heap out of the GC's sight
        import gc
        gc.collect()
        gc.freeze()
    e2e_steps = max(1, min(args.steps, 5))
    page_s, tab = ctx.wall(lambda: api_call(adata_page), 1, e2e_steps)
    assert np.array_equal(tab.values, ref_tab), "API table (pageable) differs from the device-resident run"
    Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
    Xh.copy_(Xd)
    torch.cuda.synchronize()
    adata_pin = _adata(Xh.numpy(), s_lab, c_lab, G)
    e2e_s, tab = ctx.wall(lambda: api_call(adata_pin), max(1, min(args.warmup, 2)), e2e_steps)
    e2e_calls_ms = list(ctx.last_wall_ms)
    if os.environ.get("MN_BENCH_TRACE"):      # diagnostic: device timeline of a few more API calls (rank 0 prints)
        orig = _lib.call
        for rep in range(int(os.environ["MN_BENCH_TRACE"])):
            log = []

            def traced(name, *a):
                t = time.perf_counter()
                orig(name, *a)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                log.append((name, t, ev))
            _lib.call = traced
            torch.cuda.synchronize()
            ev0 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            api_call(adata_pin)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            _lib.call = orig
            if rank == 0 and engine.UPLOAD_LOG:
                sys.stderr.write("   upload chunks (rows, host enqueue ms, device done ms): " + " ".join(
                    f"[{lo // 1000}k {1e3 * (th - t0):.1f} {ev0.elapsed_time(ev):.1f}]" for lo, hi, ev, th in engine.UPLOAD_LOG) + "\n")
            engine.UPLOAD_LOG.clear()
            if rank == 0:
                sys.stderr.write(f"traced e2e call {rep}: wall {1e3 * (t1 - t0):.1f} ms\n")
                for name, t, ev in log:
                    sys.stderr.write(f"   {name:24s} host +{1e3 * (t - t0):7.2f} ms   device done +{ev0.elapsed_time(ev):7.2f} ms\n")
    # (N > 1: every rank uploads its own 1/N slice of the matrix -- the total over ranks is the matrix once --
    #  plus the small label arrays on every rank)
    h2d = int(N) * int(G) * 4 + world * (int(N) * 4 * 3 + (S + 1) * 4)
    d2h = L * L * 8 + L * 4 + int(N)
    del Xh, adata_pin, adata_page, X_page

    # ---- CPU baseline (rank 0, N = 1) + parity -----------------------------------
    cpu = None
    oracle_tab = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, ncell, otab = cpu_reference_time(G, K, REF_SAMPLE_CELLS_PER_STUDY, S)
        oracle_tab = otab
        cpu = {"value": v, "unit": UNIT, "cores": _blas_threads(), "kind": "port",
               "sample": f"oracle port (NumPy float64) on {S}x{REF_SAMPLE_CELLS_PER_STUDY} cells x {G} genes of the same "
                         f"generator, {dt:.1f} s for N^2 = {ncell * ncell:.2e} pairs; the reference's cost grows like "
                         f"N^2 log N and its N x N float64 matrix (320 GB at full size) does not fit the host"}
    parity = parity_block(ctx, oracle_tab)

    # ---- the other BASELINE configs ------------------------------------------------
    extra = {}
    if not args.no_extras and args.config == "c2":
        del Xd
        engine.release_spill_cache()
        for name, fn in (("c1", lambda: bench_fast_config(ctx, "c1")), ("c4", lambda: bench_fast_config(ctx, "c4")),
                         ("c5", lambda: bench_c5(ctx)), ("c3", lambda: bench_c3(ctx))):
            try:
                extra[name] = fn()
            except Exception as e:               # a failing side config must not take the headline line with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = ctx.peaks
        flops_hist = (N * (N - 1) / 2.0) * 2.0 * G / world
        kern = {"corr_hist": {"ms": stage_ms["corr_hist"], "flop": flops_hist},
                "corr_vote": {"ms": stage_ms["corr_vote"], "flop": flops_hist}}
        dom = max(kern, key=lambda k: kern[k]["ms"])
        peak = peaks["tf_sust"]
        # DRAM bytes per launch measured by `ncu --set full` on this very command (profiles/, N = 1 only)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_c2.json")
        if world == 1 and args.config == "c2" and os.path.isfile(tpath):
            traffic = json.load(open(tpath))
        roof = {}
        for k, v in kern.items():
            ach = v["flop"] / (v["ms"] / 1e3) / 1e12
            roof[k] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                       "traffic": traffic.get(k), "ms_per_launch": v["ms"],
                       "flop_counted": "N(N-1)/2 * 2G per pass: the network is symmetric and only its upper triangle is computed",
                       "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['src']})",
                       "traffic_source": "profiles/ncu_traffic_c2.json (ncu --set full of this command)" if traffic else None}
        hbm = {}
        if sp.get("spill_tiles", 0) >= sp.get("total_tiles", 1):
            # pass 2 streamed every tile from the rho spill: an HBM-bound kernel
            b = float(sp["spill_tiles"]) * sp["tile_bytes"]
            ach = b / (stage_ms["corr_vote"] / 1e3) / 1e9
            roof["corr_vote"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm"], "traffic": traffic.get("corr_vote_spill"),
                                 "ms_per_launch": stage_ms["corr_vote"],
                                 "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['src']})"}
        bytes_rank = 6.0 * N * G           # K1: 4 B read + 2 B fp16 plane written per element
        ach = bytes_rank / (stage_ms["rank_normalize"] / 1e3) / 1e9
        hbm["rank_normalize"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm"], "traffic": traffic.get("rank_normalize"),
                                 "ms_per_launch": stage_ms["rank_normalize"]}
        bytes_auroc = 4.0 * N * L + 4.0 * N
        ach = bytes_auroc / (stage_ms["auroc"] / 1e3) / 1e9
        hbm["auroc"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach / peaks["hbm"], "traffic": traffic.get("auroc"), "ms_per_launch": stage_ms["auroc"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "steps_ms": steps_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 exact-integer operands, f32 TMEM accumulate, int64 vote sums",
            "data": "synthetic",
            "config": {"workload": desc, "n_cells": N, "n_genes": G, "n_labels": L, "nbins": engine.DEFAULT_NBINS,
                       "rho_spill": {"tiles": sp.get("spill_tiles", 0), "of": sp.get("total_tiles", 0),
                                     "gb": round(sp.get("spill_tiles", 0) * sp.get("tile_bytes", 0) / 1e9, 1)},
                       "l2": "inputs (1.6 GB fp32 + 0.8 GB fp16 planes) are larger than the 126 MB L2",
                       "parallelism": f"row-tile sharding x{world}" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "calls_ms": e2e_calls_ms,
                    "input": "pinned host array, warm process"},
            "e2e_pageable": {"value": pairs / page_s, "unit": UNIT, "ms_per_step": page_s * 1e3,
                             "input": "pageable NumPy array, warm process"},
            "e2e_cold": {"value": pairs / cold_s, "unit": UNIT, "ms": cold_s * 1e3,
                         "input": "pageable NumPy array, first call of the process (CUDA library load, spill and staging allocation)"},
            "gpu_launches": launches,
            "table_sha256": _sha(ref_tab),
            "parity": parity,
            "roofline": roof[dom],
            "kernels": {**roof, **hbm},
            "stage_ms": stage_ms,
            "configs": extra,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        engine.release_peer_buffers()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=list(CONFIGS), default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `configs` block (c1, c3, c4, c5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
