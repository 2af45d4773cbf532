#!/usr/bin/env python
"""variableGenes on the GPU vs the NumPy restatement of the reference.
usage: python tools/bench_vargenes.py [cells_per_study] [genes] [studies]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
Xd, study, _ = make_counts_torch(2, S, n, G, 10, dev)
N = Xd.shape[0]
perm = np.argsort(study, kind="stable").astype(np.int32)
seg = np.concatenate([[0], np.cumsum(np.bincount(study, minlength=S))]).astype(np.int32)
perm_d, seg_d = engine._i32(perm), engine._i32(seg)
med = torch.empty((S, G), dtype=torch.float64, device=dev)
var = torch.empty((S, G), dtype=torch.float64, device=dev)


def run():
    engine._lib.call("mn_gene_stats", Xd, int(Xd.stride(0)), perm_d, seg_d, S, G, med, var)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
alg = 4.0 * N * G                      # the matrix once (algorithmic); the radix select sweeps it 4 times
print(f"mn_gene_stats N={N} G={G} S={S}: {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic "
      f"({4 * alg / ms / 1e6:.0f} GB/s moved, 4 sweeps)")
# CPU restatement on a bounded sample: one study
x = Xd[: seg[1]].cpu().numpy()
t0 = time.perf_counter()
np.median(x, axis=0)
np.var(x, axis=0)
dt = time.perf_counter() - t0
print(f"NumPy median + var of one study ({x.shape[0]} cells): {dt * 1e3:.0f} ms -> {S * dt * 1e3:.0f} ms for {S} studies")
