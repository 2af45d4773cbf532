#!/usr/bin/env python
"""Timeline of the end-to-end call at config 2 (pinned host matrix): for every C-ABI call of one
pymn_b200.MetaNeighborUS invocation, when the host enqueued it and when the device finished it (CUDA events),
both relative to the start of the call.  Shows where the device idles.  usage: python tools/trace_e2e.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

import pymn_b200
from pymn_b200 import _lib
from pymn_b200.anndata_lite import AnnDataLite
from pymn_b200.synth import make_counts_torch, make_labels

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 5
S, n, G, K = 4, 50000, 2000, 10
dev = torch.device("cuda", 0)
Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
Xh = torch.empty(Xd.shape, dtype=torch.float32, pin_memory=True)
Xh.copy_(Xd)
del Xd
s, c, _, _ = make_labels(S, n, K)
obs = pd.DataFrame({"study_id": s, "cell_type": c})
var = pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=pd.Index([f"g{i}" for i in range(G)], dtype=object))
ad = AnnDataLite(Xh.numpy(), obs, var)

log = []
orig = _lib.call


def traced(name, *args):
    t = time.perf_counter()
    orig(name, *args)
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    log.append((name, t, ev))


_lib.call = traced
import pymn_b200.engine as E
E._lib.call = traced


def call():
    return pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=False, symmetric_output=False, save_uns=False)


for _ in range(2):
    call()
for k in range(n_calls):
    log.clear()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    call()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"call {k}: wall {1e3 * (t1 - t0):.1f} ms")
    if k in (0, n_calls - 1):
        prev = 0.0
        for name, t, ev in log:
            done = e0.elapsed_time(ev)
            print(f"   {name:24s} host +{1e3 * (t - t0):7.2f} ms   device done +{done:7.2f} ms   (+{done - prev:6.2f})")
            prev = done
