#!/usr/bin/env python
"""bench.py -- MetaNeighborUS cell-pairs/s on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[1], the config the metric is quoted on and the
largest that fits one GPU): MetaNeighborUS fast_version=False, full Spearman
cell x cell network, 4 studies x 50,000 cells x 2,000 HVGs, 40 clusters
(N = 200,000; N^2 = 4e10 cell pairs), synthetic counts (pymn_b200/synth.py).

A "step" is one pass of the hot path over that input: per-cell rank + normalise
-> pass 1 (rho histogram on tensor cores) -> CDF LUT -> pass 2 (rank-transformed
votes + node degree) -> vote matrix -> rank-sum AUROC table.  `value` = N^2 /
step time with inputs resident in HBM; `e2e` = the same metric through the public
API (pymn_b200.MetaNeighborUS on an AnnData-like object whose X is a pinned host
array: label encoding, H2D of X, all kernels, D2H of the table, DataFrame) .

N > 1 (torchrun, one process per GPU): the two tensor-core passes are split by
128-row tile across ranks, integer partials are summed with NCCL all-reduce
(strong scaling: the workload is fixed).
"""
from __future__ import annotations

import argparse
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
REF_SAMPLE_CELLS_PER_STUDY = 1000        # CPU arm: 4 x 1,000 cells of the same generator (N^2 = 1.6e7 pairs)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 100 ms via NVML while running."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:          # NVML unavailable: report what we have
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _labels(S, n, K):
    from pymn_b200.synth import make_labels
    s, c, _, _ = make_labels(S, n, K)
    return s, c


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on host cores
# ---------------------------------------------------------------------------
def cpu_reference_time(G, K, n_per_study, S=4, seed=2, repeats=1):
    """Times oracle.mn_oracle.metaneighbor_us_default (NumPy float64 restatement of
    MetaNeighborUS.py:145-224 + utils.py:35-75) on a bounded sample.  Returns (pairs/s, seconds, N)."""
    from oracle import mn_oracle as O
    from pymn_b200.synth import make_counts
    x, s, c = make_counts(seed, S, n_per_study, G, K)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.metaneighbor_us_default(x, s, c, True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n = x.shape[0]
    return n * n / best, best, n


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
        v, dt, ncell = cpu_reference_time(G, K, REF_SAMPLE_CELLS_PER_STUDY, S)
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
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pymn_b200
    from pymn_b200 import _lib, engine
    from pymn_b200.anndata_lite import AnnDataLite
    from pymn_b200.synth import make_counts_torch
    from pymn_b200.utils import label_layout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA GPUs"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"

    S, n, G, K, desc = CONFIGS[args.config]
    N = S * n
    pairs = float(N) * float(N)

    # synthetic inputs, identical on every rank (same seed)
    Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
    s_lab, c_lab = _labels(S, n, K)
    lay = label_layout(s_lab, c_lab)
    L = len(lay.row_labels)
    perm, cell_label, seg_off = engine._i32(lay.perm), engine._i32(lay.cell_label), engine._i32(lay.seg_off)
    shard = engine.dist_shard()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(timer=None):
        return engine.us_default_device(Xd, perm, cell_label, seg_off, L, S, True, engine.DEFAULT_NBINS, shard, timer)

    # ---- device-resident timing ------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    timer = engine.StageTimer()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(timer)
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_step = float(t.item()) / args.steps
    value = pairs / (ms_step / 1e3)
    stage_ms = timer.mean_ms()
    auc_dev = out[0]
    assert bool(torch.isfinite(auc_dev).all()), "AUROC table has non-finite entries"

    # ---- end to end through the public API, host buffers -------------------------
    import pandas as pd
    Xh = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
    Xh.copy_(Xd)
    torch.cuda.synchronize()
    obs = pd.DataFrame({"study_id": s_lab, "cell_type": c_lab})
    var = pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=pd.Index([f"g{i}" for i in range(G)], dtype=object))
    adata = AnnDataLite(Xh.numpy(), obs, var)

    class _AllGenes(AnnDataLite):
        """adata[:, all genes] would copy 1.6 GB on the host just to select every column:
        the bench's AnnData returns itself for the identity gene selection."""
        def __getitem__(self, key):
            return self
    adata.__class__ = _AllGenes

    def e2e_step():
        return pymn_b200.MetaNeighborUS(adata, "study_id", "cell_type", fast_version=False, symmetric_output=False,
                                        save_uns=False)
    e2e_warm = max(1, min(args.warmup, 2))
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tab = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    # (N > 1: every rank uploads its own 1/N slice of the matrix -- the total over ranks is the matrix once --
    #  plus the small label arrays on every rank)
    h2d = int(N) * int(G) * 4 + world * (int(N) * 4 * 3 + (S + 1) * 4)
    d2h = L * L * 8 + L * 4 + int(N)
    # the API result must equal the device-resident result (same kernels, same inputs)
    order = np.argsort(lay.sorted_pos)
    ref_tab = auc_dev.cpu().numpy()[np.ix_(order, order)].T
    assert np.array_equal(tab.values, ref_tab), "API table differs from the device-resident run"

    if rank == 0:
        peaks = _peaks()
        flops_hist = (N * (N - 1) / 2.0) * 2.0 * G / world
        flops_vote = flops_hist            # pass 2 also covers the upper triangle only (symmetric network)
        kern = {
            "corr_hist": {"ms": stage_ms["corr_hist"], "flop": flops_hist},
            "corr_vote": {"ms": stage_ms["corr_vote"], "flop": flops_vote},
        }
        dom = max(kern, key=lambda k: kern[k]["ms"])
        sp = dict(engine.LAST_SPILL)
        peak = peaks["tf_sust"]
        # DRAM bytes per launch measured by `ncu --set full` on this very command (profiles/, N=1 only)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_c2.json")
        if world == 1 and args.config == "c2" and os.path.isfile(tpath):
            traffic = json.load(open(tpath))
        roof = {}
        for k, v in kern.items():
            ach = v["flop"] / (v["ms"] / 1e3) / 1e12
            roof[k] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                       "traffic": traffic.get(k), "ms_per_launch": v["ms"],
                       "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['src']})"}
        hbm = {}
        if sp.get("spill_tiles", 0) >= sp.get("total_tiles", 1):
            # pass 2 streamed every tile from the rho spill: an HBM-bound kernel, 4 B per cell pair of the
            # upper triangle (tile granularity: spill_tiles x 128 x 256 x 4 B)
            b = float(sp["spill_tiles"]) * sp["tile_bytes"]
            ach = b / (stage_ms["corr_vote"] / 1e3) / 1e9
            roof["corr_vote"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm"], "traffic": traffic.get("corr_vote_spill"),
                                 "ms_per_launch": stage_ms["corr_vote"],
                                 "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['src']})"}
        bytes_rank = 6.0 * N * G           # K1: 4 B read + 2 B fp16 plane written per element
        ach = bytes_rank / (stage_ms["rank_normalize"] / 1e3) / 1e9
        hbm["rank_normalize"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm"], "traffic": traffic.get("rank_normalize"), "ms_per_launch": stage_ms["rank_normalize"]}
        bytes_auroc = 4.0 * N * L + 4.0 * N
        ach = bytes_auroc / (stage_ms["auroc"] / 1e3) / 1e9
        hbm["auroc"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach / peaks["hbm"], "traffic": traffic.get("auroc"), "ms_per_launch": stage_ms["auroc"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 exact-integer operands, f32 TMEM accumulate, int64 vote sums",
            "data": "synthetic",
            "config": {"workload": desc, "n_cells": N, "n_genes": G, "n_labels": L, "nbins": engine.DEFAULT_NBINS,
                       "rho_spill": {"tiles": sp.get("spill_tiles", 0), "of": sp.get("total_tiles", 0),
                                     "gb": round(sp.get("spill_tiles", 0) * sp.get("tile_bytes", 0) / 1e9, 1)},
                       "l2": "inputs (1.6 GB fp32 + 0.8 GB fp16 planes) are larger than the 126 MB L2",
                       "parallelism": f"row-tile sharding x{world}" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
            "gpu_launches": int(lt.item()),
            "roofline": roof[dom],
            "kernels": {**roof, **hbm},
            "stage_ms": stage_ms,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, ncell = cpu_reference_time(G, K, REF_SAMPLE_CELLS_PER_STUDY, S)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": _blas_threads(), "kind": "port",
                "sample": f"oracle port (NumPy float64) on {S}x{REF_SAMPLE_CELLS_PER_STUDY} cells x {G} genes of the same "
                          f"generator, {dt:.1f} s for N^2 = {ncell * ncell:.2e} pairs; the reference's cost grows like "
                          f"N^2 log N and its N x N float64 matrix (320 GB at full size) does not fit the host"}
        print(json.dumps(line), flush=True)
    if world > 1:
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
