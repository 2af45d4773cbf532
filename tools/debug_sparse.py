"""GPU debug: full-network path on 8 sparse genes vs the host restatement (tests/lut_emulation.py), stage by stage."""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from oracle import mn_oracle as O
from tests import lut_emulation as E
from pymn_b200 import engine
from pymn_b200.utils import label_layout
from pymn_b200.synth import make_counts

x, s, c = make_counts(3, 3, 200, 400, 5)
for g in (8, 12):
    rng = np.random.default_rng(g)
    sub = np.ascontiguousarray(x[:, rng.choice(400, size=g, replace=False)])
    lay = label_layout(s, c)
    perm = np.asarray(lay.perm)
    xs = np.ascontiguousarray(sub[perm])
    rc = engine.rank_normalize(engine.to_device(xs))
    L = len(lay.row_labels)
    lab = np.asarray(lay.cell_label).astype(np.int32)
    nb = engine.DEFAULT_NBINS
    votes, degree, hist, lut = engine.corr_network_votes(rc, engine._i32(lab), L, nbins=nb)
    torch.cuda.synchronize()
    d, inv = E.planes(xs)
    inv_g = rc.inv_norm.cpu().numpy()
    print(f"g={g}: inv_norm equal: {np.array_equal(inv, inv_g)}  dropped {int((inv_g == 0).sum())}")
    q = E.q_matrix(d, inv_g, nb)
    kept = inv_g > 0
    rh = E.hist_upper(q, kept, nb)
    gh = hist.cpu().numpy()
    print("  hist equal:", np.array_equal(rh, gh), " max cum diff", np.abs(np.cumsum(gh - rh)).max(), " n bins differing", int((rh != gh).sum()))
    lut_h = lut.cpu().numpy().astype(np.int64)
    print("  lut equal (from GPU hist):", np.array_equal(lut_h, E.build_lut(gh, len(d))), " heavy bins", lut_h[nb + 1])
    w = np.triu(E.w_from_lut(lut_h, q), 1); w = w + w.T; w[~kept] = 0; w[:, ~kept] = 0; np.fill_diagonal(w, 0)
    ref_votes = np.stack([w[:, lab == l].sum(axis=1) for l in range(L)], axis=1)
    gv = votes.cpu().numpy(); gd = degree.cpu().numpy()
    dv = np.abs(gv - ref_votes)
    print("  votes: max abs diff", dv.max(), "of", ref_votes.max(), " rows differing", int((dv.max(axis=1) > 0).sum()), " degree max diff", np.abs(gd - w.sum(axis=1)).max())
    if dv.max() > 0:
        i = int(np.argmax(dv.max(axis=1)))
        # which pairs of row i differ?  recompute row i's weights through the GPU LUT with GPU-like orientation
        print("  worst row", i, "label", lab[i], "kept", kept[i], "diff per label", (gv[i] - ref_votes[i]))
    # AUROC from the GPU's own integer votes through the oracle arithmetic vs the GPU's AUROC
    r = engine.us_default(sub, lay, True)
    order = np.argsort(lay.sorted_pos)
    got = r["auroc"][np.ix_(order, order)].T
    emu, _ = E.us_default_from_w(E.weights(sub, nb), s, c)
    print("  auroc GPU vs emu max", np.nanmax(np.abs(got - emu)))
    wf = w.astype(np.float64) / E.W_ONE; np.fill_diagonal(wf, 1.0)
    # same weights, sorted order
    s_s, c_s = np.asarray(s, dtype=object)[perm], np.asarray(c, dtype=object)[perm]
    emu2, _ = E.us_default_from_w(wf, s_s, c_s)
    print("  auroc emu(sorted order, GPU lut) vs emu(original order)", np.nanmax(np.abs(emu2 - emu)))
    wg = np.zeros_like(wf)
    # GPU integer votes -> float votes as votes_finalize does, then oracle AUROC
    one = float(E.W_ONE)
    v = gv.astype(np.float64).copy(); v[np.arange(len(lab)), lab] += one
    vm = (v / (gd.astype(np.float64) + one)[:, None]).astype(np.float32)
    out = np.full((L, L), np.nan)
    seg = np.asarray(lay.seg_off)
    for si in range(len(seg) - 1):
        rows = np.arange(seg[si], seg[si + 1])
        ranks = O._nanrank_cols(vm[rows].astype(np.float64))
        for ct in np.unique(lab[rows]):
            pos = lab[rows] == ct
            n_p = float(pos.sum()); nn = len(rows) - n_p
            out[:, ct] = (ranks[pos].sum(axis=0) / n_p - (n_p + 1) / 2) / nn
    gA = r["auroc"]       # rows = test labels (row order), cols = train
    print("  GPU auroc vs oracle-AUROC of GPU votes (float32):", np.nanmax(np.abs(gA - out.T)))
    v64 = (v / (gd.astype(np.float64) + one)[:, None])
    out64 = np.full((L, L), np.nan)
    for si in range(len(seg) - 1):
        rows = np.arange(seg[si], seg[si + 1])
        ranks = O._nanrank_cols(v64[rows])
        for ct in np.unique(lab[rows]):
            pos = lab[rows] == ct
            n_p = float(pos.sum()); nn = len(rows) - n_p
            out64[:, ct] = (ranks[pos].sum(axis=0) / n_p - (n_p + 1) / 2) / nn
    print("  float32 votes vs float64 votes (same integers) AUROC diff:", np.nanmax(np.abs(out - out64)))
