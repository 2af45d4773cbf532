#!/usr/bin/env python
"""One device-resident step of the headline workload (config 2, or a fraction of its cells) and nothing else:
the shortest program that launches the final pass-1 / pass-2 kernels at full size, for `ncu --replay-mode
application` (kernel replay cannot save / restore the 80 GB rho spill).  usage: python tools/prof_c2.py [cells/study] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch, make_labels
from pymn_b200.utils import label_layout

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S, G, K = 4, 2000, 10
dev = torch.device("cuda", 0)
Xd, _, _ = make_counts_torch(2, S, n, G, K, dev)
s, c, _, _ = make_labels(S, n, K)
lay = label_layout(s, c)
L = len(lay.row_labels)
args = (Xd, engine._i32(lay.perm), engine._i32(lay.cell_label), engine._i32(lay.seg_off), L, S, True, engine.DEFAULT_NBINS)
for _ in range(steps):
    t = engine.StageTimer()
    auc = engine.us_default_device(*args, timer=t)[0]
    torch.cuda.synchronize()
    print({k: round(v, 3) for k, v in t.mean_ms().items()}, "mean AUROC", float(auc.mean()))
