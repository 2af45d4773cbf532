#!/usr/bin/env python
"""Supervised MetaNeighbor throughput (BASELINE config 3): gene-sets/s through the public API.

usage: python tools/bench_supervised.py [n_sets] [--fast]
  20,000 cells x 3 studies x 5,000-gene universe, 10 shared cell types, gene sets of 10-200 genes
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

import pymn_b200
from pymn_b200.anndata_lite import AnnDataLite
from pymn_b200.synth import make_counts_torch, make_genesets, make_labels


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_sets = int(args[0]) if args else 100
    fast = "--fast" in sys.argv
    S, n, G, K = 3, 6667, 5000, 10
    dev = torch.device("cuda", 0)
    Xd, _, _ = make_counts_torch(3, S, n, G, K, dev)
    X = Xd.cpu().numpy()
    s, c, _, _ = make_labels(S, n, K)
    genes = pd.Index([f"g{i}" for i in range(G)], dtype=object)
    ad = AnnDataLite(X, pd.DataFrame({"study_id": s, "cell_type": c}), pd.DataFrame(index=genes))
    gs = pd.DataFrame(make_genesets(3, G, n_sets, 10, 200), index=genes, columns=[f"set{j}" for j in range(n_sets)])
    pymn_b200.MetaNeighbor(ad, "study_id", "cell_type", gs.iloc[:, :2], fast_version=fast, save_uns=False)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tab = pymn_b200.MetaNeighbor(ad, "study_id", "cell_type", gs, fast_version=fast, save_uns=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"supervised fast_version={fast}: {n_sets} gene sets over N={X.shape[0]} cells in {dt:.2f} s "
          f"= {n_sets / dt:.1f} gene-sets/s (incl. H2D of the {X.nbytes / 1e6:.0f} MB matrix); "
          f"mean AUROC {np.nanmean(tab.values):.3f}")


if __name__ == "__main__":
    main()
