#!/usr/bin/env python
"""Where the supervised path's time goes: host enqueue time vs device time per gene set (config 3 shape).
usage: python tools/prof_supervised.py [n_sets] [--fast]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pymn_b200 import engine
from pymn_b200.synth import make_counts_torch, make_genesets, make_labels
from pymn_b200.utils import joint_layout


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_sets = int(args[0]) if args else 100
    fast = "--fast" in sys.argv
    S, n, G, K = 3, 6667, 5000, 10
    dev = torch.device("cuda", 0)
    Xd, _, _ = make_counts_torch(3, S, n, G, K, dev)
    s, c, _, _ = make_labels(S, n, K)
    lay, ct = joint_layout(s, c)
    gm = make_genesets(3, G, n_sets, 10, 200)
    cols = [np.flatnonzero(gm[:, j]).astype(np.int32) for j in range(n_sets)]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in cols])]).astype(np.int64)
    all_cols = engine._i32(np.concatenate(cols))
    ctx = engine.supervised_prepare(Xd, lay)
    engine.supervised_tables(ctx, all_cols, offs[:3], K, True, fast)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    tables, _ = engine.supervised_tables(ctx, all_cols, offs, K, True, fast)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"fast={fast} n_sets={n_sets}: host enqueue {1e3 * t_host / n_sets:.3f} ms/set, device span "
          f"{e0.elapsed_time(e1) / n_sets:.3f} ms/set, wall {1e3 * t_all / n_sets:.3f} ms/set "
          f"= {n_sets / t_all:.0f} sets/s; mean AUROC {float(torch.nanmean(tables)):.4f}")


if __name__ == "__main__":
    main()
