#!/usr/bin/env python
"""Micro-benchmark of K1 (rank_normalize) alone: CUDA-event time and achieved GB/s, fp16-plane output
(full-network path) and int8-limb output (centroid paths).
usage: python tools/bench_k1.py [n_cells] [n_genes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
g = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda", 0)
X, _, _ = make_counts_torch(2, 4, n // 4, g, 10, dev)
nnz = float((X != 0).float().mean().item())
mx = float(X.max().item())
for limbs in (False, True):
    out_b = 2.0 if limbs else (2.0 if g <= 2048 else 4.0)
    for _ in range(2):
        rc = engine.rank_normalize(X, limbs=limbs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rc = engine.rank_normalize(X, limbs=limbs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"K1 n={X.shape[0]} g={g} nnz_frac={nnz:.3f} max={mx:.0f} out={'int8 limbs' if limbs else 'fp16 planes'}: {ms:.3f} ms  "
          f"{(4.0 + out_b) * X.shape[0] * g / ms / 1e6:.1f} GB/s ({4 + out_b:.0f} B/element moved)")
    del rc
