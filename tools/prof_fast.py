#!/usr/bin/env python
"""The centroid path through engine.us_fast_device at a BASELINE shape, for ncu launch lists / profiles.
usage: python tools/prof_fast.py c1|c4 [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch, make_labels
from pymn_b200.utils import label_layout

PRESETS = {"c1": (2, 10000, 2000, 20, False), "c4": (7, 71429, 3000, 150, True), "c4s": (7, 20000, 3000, 150, True)}
S, n, G, K, ovb = PRESETS[sys.argv[1]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
Xd, _, _ = make_counts_torch(4, S, n, G, K, dev)
s, c, _, _ = make_labels(S, n, K)
lay = label_layout(s, c)
for r in range(reps):
    t = engine.StageTimer()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = engine.us_fast_device(Xd, lay, True, ovb, t)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"rep {r}: wall {wall:.2f} ms  stages {{" + ", ".join(f"{k}: {v:.3f}" for k, v in t.mean_ms().items()) + "}")
