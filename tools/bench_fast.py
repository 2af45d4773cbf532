#!/usr/bin/env python
"""Stage timings of the centroid ("fast_version") path at BASELINE shapes (device-resident inputs).

usage: python tools/bench_fast.py c1|c4|c5|<S> <cells/study> <genes> <types/study> [--ovb] [--stages]
  c1: 2 x 10,000 x 2,000 x 20      c4: 7 x 71,429 x 3,000 x 150 (one_vs_best)
  c5: pretrained: 150-cluster model scored against 1,000,000 query cells x 3,000 genes (50 query labels)
Under torchrun (one process per GPU) the train columns are sharded across the ranks
(engine._score_against_centroids); rank 0 prints the max-over-ranks step time.
--stages times K3 / K6 / K7 separately (single GPU only).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch, make_labels
from pymn_b200.utils import label_layout

PRESETS = {"c1": (2, 10000, 2000, 20, False), "c4": (7, 71429, 3000, 150, True), "c5": (1, 1000000, 3000, 50, False)}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ovb = "--ovb" in sys.argv
    stages = "--stages" in sys.argv
    pre = args[0] if args and args[0] in PRESETS else None
    if pre:
        S, n, G, K, ovb0 = PRESETS[pre]
        ovb = ovb or ovb0
    else:
        S, n, G, K = (int(v) for v in args[:4])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    Xd, _, _ = make_counts_torch(4, S, n, G, K, dev)
    s, c, _, _ = make_labels(S, n, K)
    t0 = time.perf_counter()
    lay = label_layout(s, c)
    t_host = time.perf_counter() - t0
    L, N = len(lay.row_labels), Xd.shape[0]
    pretrained = pre == "c5"
    Lt, St = 150, 3
    if pretrained:
        rng = np.random.default_rng(0)
        model_c = torch.as_tensor(rng.normal(size=(Lt, G)).astype(np.float32) * 30, device=dev)
        model_n = torch.as_tensor(rng.integers(200, 5000, size=Lt).astype(np.float32), device=dev)
        model_study = engine._i32(np.arange(Lt) % St)
    perm, lab_row_off = engine._i32(lay.perm), engine._i32(lay.lab_row_off)
    train_study = engine._i32(lay.label_study)

    def step(tm):
        tm.start("rank_normalize")
        rc = engine.rank_normalize(Xd, row_index=perm, drop_mode=0)
        tm.stop("rank_normalize")
        tm.start("centroids")
        if pretrained:
            C = torch.zeros((Lt + St, rc.ldd), dtype=torch.float32, device=dev)
            C[:Lt, :G] = model_c
            nv = torch.zeros((Lt + St,), dtype=torch.float32, device=dev)
            nv[:Lt] = model_n
            engine.group_sum_into(C, nv, Lt, model_study, St, G)
            n_train, tr_study, n_st = Lt, model_study, St
        else:
            C, nv = engine.centroids(rc, lab_row_off, L, extra_rows=S)
            engine.group_sum_into(C, nv, L, train_study, S, G)
            n_train, tr_study, n_st = L, train_study, S
        tm.stop("centroids")
        if stages and world == 1:
            n_b = n_train + n_st
            b_hi, b_lo, inv_scale = engine.split_f16(C[:n_b], G)
            tm.start("votes_gemm")
            out = engine.votes_fast(rc, b_hi, b_lo, inv_scale, nv[:n_b].contiguous(), n_b)
            tm.stop("votes_gemm")
            cl = engine._i32(lay.cell_label)
            cell_label = torch.where(rc.inv_norm > 0, cl, torch.full_like(cl, -1))
            tm.start("auroc")
            auc, _ = engine.auroc(out, n_train, out[n_train:], tr_study, engine._i32(lay.seg_off), len(lay.studies),
                                  cell_label, L)
            tm.stop("auroc")
            if ovb:
                tm.start("one_vs_best")
                engine.one_vs_best(out, n_train, out[n_train:], tr_study, engine._i32(lay.seg_lab_off),
                                   len(lay.studies), lab_row_off, cell_label, L, auc)
                tm.stop("one_vs_best")
            return auc
        tm.start("votes+auroc" + ("+1vBest" if ovb else ""))
        res, auc, _ = engine._score_against_centroids(rc, lay, C, nv, n_train, tr_study, n_st, True, ovb)
        tm.stop("votes+auroc" + ("+1vBest" if ovb else ""))
        return auc

    for _ in range(2):
        step(engine.StageTimer())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = engine.StageTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        auc = step(tm)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = tm.mean_ms()
    if rank == 0:
        n_train = Lt if pretrained else L
        gb = {"rank_normalize": (4.0 + (4.0 if G > 2048 else 2.0)) * N * G / 1e9, "auroc": (4.0 * N * n_train + 4.0 * N) / 1e9}
        print(f"fast path S={S} N={N} G={G} L={L} ovb={ovb} pretrained={pretrained} gpus={world}: "
              f"{float(ms.item()):.2f} ms/step (host label layout {t_host * 1e3:.0f} ms)")
        for k, v in st.items():
            extra = f"  {gb[k] / v * 1e3:8.0f} GB/s" if k in gb else ""
            if k == "votes_gemm":
                extra = f"  {2.0 * N * G * (n_train + 3) / v / 1e9:8.1f} TFLOP/s"
            print(f"   {k:22s} {v:9.3f} ms{extra}")
        print("   auroc finite:", bool(torch.isfinite(auc).all().item()), " mean:", float(auc.mean().item()))
    if "--e2e" in sys.argv and not pretrained:
        # end to end through the public API: pinned host matrix -> pymn_b200.MetaNeighborUS -> host table
        # (label encoding, upload -- each rank only its slice of the cells on N > 1 GPUs --, kernels, D2H)
        import pandas as pd
        import pymn_b200
        from pymn_b200.anndata_lite import AnnDataLite
        Xh = torch.empty(tuple(Xd.shape), dtype=torch.float32, pin_memory=True)
        Xh.copy_(Xd)
        torch.cuda.synchronize()
        var = pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=pd.Index([f"g{i}" for i in range(G)], dtype=object))

        class _AllGenes(AnnDataLite):
            def __getitem__(self, key):          # the identity gene selection would copy the host matrix
                return self
        ad = _AllGenes(Xh.numpy(), pd.DataFrame({"study_id": s, "cell_type": c}), var)

        def call():
            t = pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=True, one_vs_best=ovb,
                                         symmetric_output=False, save_uns=False)
            torch.cuda.synchronize()
            return t
        call()
        if world > 1:
            dist.barrier()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        te = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"   e2e through pymn_b200.MetaNeighborUS (pinned host matrix, {Xh.numel() * 4 / 1e9:.1f} GB): "
                  f"{float(te.item()) * 1e3:.0f} ms per call on {world} GPU(s)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
