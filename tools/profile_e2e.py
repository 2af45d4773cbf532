#!/usr/bin/env python
"""Where does the end-to-end time of pymn_b200.MetaNeighborUS (config 2, pinned host matrix) go?
cProfile of one call after warm-up; usage: python tools/profile_e2e.py [cells_per_study]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

import pymn_b200
from pymn_b200.anndata_lite import AnnDataLite
from pymn_b200.synth import make_counts_torch, make_labels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
S, G, K = 4, 2000, 10
dev = torch.device("cuda", 0)
Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
Xh = torch.empty(Xd.shape, dtype=torch.float32, pin_memory=True)
Xh.copy_(Xd)
del Xd
s, c, _, _ = make_labels(S, n, K)
obs = pd.DataFrame({"study_id": s, "cell_type": c})
var = pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=pd.Index([f"g{i}" for i in range(G)], dtype=object))


ad = AnnDataLite(Xh.numpy(), obs, var)


def call():
    t = pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=False, symmetric_output=False, save_uns=False)
    torch.cuda.synchronize()
    return t


for _ in range(2):
    call()
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    call()
    ts.append((time.perf_counter() - t0) * 1e3)
print("e2e calls (ms):", " ".join(f"{t:.1f}" for t in ts), f"| median {np.median(ts):.1f}")
cProfile.run("call()", "/tmp/e2e.prof")
pstats.Stats("/tmp/e2e.prof").sort_stats("cumtime").print_stats(22)
