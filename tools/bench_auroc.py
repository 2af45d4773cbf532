"""K6 alone on synthetic votes of a given shape (tools; not part of bench.py).
   python tools/bench_auroc.py [ncols] [n_seg] [seg_len] [labels_per_seg] [reps] [dist]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pymn_b200 import engine, _lib

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 150
n_seg = int(sys.argv[2]) if len(sys.argv) > 2 else 7
seg_len = int(sys.argv[3]) if len(sys.argv) > 3 else 71429
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 150
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dist = sys.argv[6] if len(sys.argv) > 6 else "gamma"
M = n_seg * seg_len
g = torch.Generator(device="cuda"); g.manual_seed(1)
if dist == "gamma":      # long right tail, like node-degree-normalised centroid votes (tools/sim in DESIGN)
    u = torch.rand((ncols + n_seg, M), generator=g, device="cuda")
    v = torch.rand((ncols + n_seg, M), generator=g, device="cuda")
    votes = 50.0 - 8.0 * torch.log(u) - 3.0 * torch.log(v)
else:
    votes = torch.randn((ncols + n_seg, M), generator=g, device="cuda")
votes[ncols:] = 5000.0 + 100.0 * torch.randn((n_seg, M), generator=g, device="cuda")
lab = (torch.arange(M, device="cuda") % seg_len * nl // seg_len + torch.arange(M, device="cuda") // seg_len * nl).to(torch.int32)
seg_off = torch.arange(0, M + 1, seg_len, dtype=torch.int32, device="cuda")
col_denom = (torch.arange(ncols, device="cuda") % n_seg).to(torch.int32)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
for r in range(2):
    engine.auroc(votes, ncols, votes[ncols:], col_denom, seg_off, n_seg, lab, n_seg * nl)
torch.cuda.synchronize()
ts = []
for r in range(reps):
    ev[2 * r].record()
    auc, _ = engine.auroc(votes, ncols, votes[ncols:], col_denom, seg_off, n_seg, lab, n_seg * nl)
    ev[2 * r + 1].record()
torch.cuda.synchronize()
ts = [ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(reps)]
ms = float(np.median(ts))
nv = ncols * M
print(f"auroc ncols={ncols} n_seg={n_seg} seg_len={seg_len} nl={nl} dist={dist}: {ms:.3f} ms  {nv / ms * 1e-6:.1f} G votes/s  "
      f"{nv * 4 / ms * 1e-6:.1f} GB/s algorithmic  mean auroc {float(torch.nanmean(auc)):.4f}")
