#!/usr/bin/env python
"""Multi-GPU outlier hunt: many device-resident steps of the full-network pipeline under torchrun, per-step wall
and device times with the collectives timed as stages of their own.  usage: torchrun ... tools/stress_dist.py [cells/study] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch, make_labels
from pymn_b200.utils import label_layout

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
S, G, K = 4, 2000, 10
Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
s, c, _, _ = make_labels(S, n, K)
lay = label_layout(s, c)
L = len(lay.row_labels)
args = (Xd, engine._i32(lay.perm), engine._i32(lay.cell_label), engine._i32(lay.seg_off), L, S, True, engine.DEFAULT_NBINS)
r, w, red = engine.dist_shard()
tm = {}


def timed_reduce(tns):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    red(tns)
    e1.record()
    tm.setdefault("allreduce_%dKB" % (tns.numel() * tns.element_size() // 1024), []).append((e0, e1))


host = []          # (what, host ms) of the instrumented calls of one iteration


def wrap(mod, name, label=None):
    f = getattr(mod, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        out = f(*a, **k)
        host.append((label or name if not (name == "call") else a[0], (time.perf_counter() - t0) * 1e3))
        return out
    setattr(mod, name, g)


from pymn_b200 import _lib
wrap(_lib, "call")
wrap(engine, "rho_spill_buffer")
wrap(engine, "_gather_columns")
wrap(engine, "auroc")
wrap(torch, "zeros")
wrap(torch, "empty")
wrap(dist, "all_reduce")
wrap(dist, "all_gather")
for _ in range(3):
    engine.us_default_device(*args, shard=(r, w, timed_reduce))
torch.cuda.synchronize()
dist.barrier()
rows = []
for it in range(iters):
    tm.clear()
    host.clear()
    t = engine.StageTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = engine.us_default_device(*args, shard=(r, w, timed_reduce), timer=t)
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = t.mean_ms()
    st.update({k: sum(a.elapsed_time(b) for a, b in v) for k, v in tm.items()})
    rows.append((wall * 1e3, t_enq * 1e3, e0.elapsed_time(e1), st, [h for h in host if h[1] > 2.0]))
walls = np.array([x[0] for x in rows])
med = float(np.median(walls))
print(f"rank {rank}: median wall {med:.2f} ms, max {walls.max():.2f} ms; outliers (> 1.3 x median):", flush=True)
for i, (wl, enq, devt, st, slow) in enumerate(rows):
    if wl > 2.5 * med:
        print(f"  rank {rank} iter {i}: wall {wl:.1f} host-enqueue {enq:.1f} device {devt:.1f} slow host calls: {slow}", flush=True)
dist.barrier()
dist.destroy_process_group()
