#!/usr/bin/env python
"""Micro-benchmark of K1 (rank_normalize) alone: CUDA-event time and achieved GB/s.
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
for _ in range(2):
    rc = engine.rank_normalize(X)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    rc = engine.rank_normalize(X)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"K1 n={X.shape[0]} g={g} nnz_frac={nnz:.3f}: {ms:.3f} ms  {6.0 * X.shape[0] * g / ms / 1e6:.1f} GB/s (6 B/element)  "
      f"mode={'cta' if os.environ.get('MN_RANK_CTA') else 'warp'}")
