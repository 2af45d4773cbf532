#!/usr/bin/env python
"""Repeatability of the node-degree GEMM (mn_votes_den_i8) at config-4 size, against the CUDA-core kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch
n, g, S = int(sys.argv[1]) if len(sys.argv) > 1 else 71429, 3000, 7
Xd, _, _ = make_counts_torch(4, S, n, g, 150, torch.device("cuda", 0))
rc = engine.rank_normalize(Xd, limbs=True)
C = torch.randn((S, rc.ldd), device="cuda") * 3.0
C[:, g:] = 0
add = torch.rand((S,), device="cuda") * 100
ref = engine.votes_den64(rc, C, add, S, cuda_cores=True)
outs = [engine.votes_den64(rc, C, add, S) for _ in range(4)]
torch.cuda.synchronize()
for i, o in enumerate(outs):
    d = (o - ref).abs().max().item()
    rel = ((o - ref).abs() / ref.abs().clamp(min=1e-30)).max().item()
    print(f"run {i}: max |tc - cuda cores| = {d:.3e} (rel {rel:.2e}); identical to run 0: {bool(torch.equal(o, outs[0]))}; "
          f"rows differing from run 0: {int((o != outs[0]).any(dim=0).sum())}")
