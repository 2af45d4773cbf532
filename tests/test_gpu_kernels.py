"""GPU parity tests, kernel by kernel, through the C ABI (pymn_b200._lib.call).

Integer / index work (2*rank, histogram counts) must be bit-exact; floating
point is held to the stated tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import cases as C
from oracle import mn_oracle as O
from tests import lut_emulation as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from pymn_b200 import engine
    return engine


def _rand_matrix(rng, n, g, kind):
    if kind == "counts":
        x = rng.poisson(rng.gamma(0.5, 2.0, size=(1, g)), size=(n, g)).astype(np.float32)
    elif kind == "dense":
        x = rng.normal(size=(n, g)).astype(np.float32)
    elif kind == "ties":
        x = rng.integers(-2, 3, size=(n, g)).astype(np.float32)
    else:
        raise ValueError(kind)
    return x


@pytest.mark.parametrize("n,g,kind", [
    (64, 257, "counts"), (33, 2000, "counts"), (17, 3000, "counts"), (40, 50, "ties"), (9, 1, "counts"),
    (21, 2, "ties"), (30, 1000, "dense"), (12, 4096, "dense"), (25, 777, "ties"), (5, 16384, "counts"),
])
def test_rank_normalize_bit_exact(eng, n, g, kind):
    rng = np.random.default_rng(n * 100003 + g)
    x = _rand_matrix(rng, n, g, kind)
    x[0] = 0.0                                   # zero-variance cell
    if n > 3:
        x[3] = 7.0                               # constant non-zero cell
    rc = eng.rank_normalize(eng.to_device(x), want_xnorm=True, want_rank2=True)
    torch.cuda.synchronize()
    ref2 = O.rank2_rows(x)
    got2 = rc.rank2.cpu().numpy()
    assert np.array_equal(got2, ref2), "2*rank differs"
    d = ref2 - (g + 1)
    hi = rc.d_hi.cpu().numpy().astype(np.float64)[:, :g]
    lo = rc.d_lo.cpu().numpy().astype(np.float64)[:, :g] if rc.d_lo is not None else 0.0
    flags = rc.flags.cpu().numpy()
    keep = (flags & 4) == 0
    assert np.array_equal((hi + lo)[keep], d[keep].astype(np.float64)), "fp16 planes are not the exact integers"
    assert np.all((hi + lo)[~keep] == 0)
    assert np.all(rc.d_hi.cpu().numpy()[:, g:] == 0), "padding columns must be zero"
    ss = (d.astype(np.float64) ** 2).sum(axis=1)
    assert np.array_equal((flags & 1) != 0, ss == 0)
    inv = rc.inv_norm.cpu().numpy()
    assert np.all(inv[~keep] == 0)
    np.testing.assert_allclose(inv[keep], 1.0 / np.sqrt(ss[keep]), rtol=1e-6)
    ref = O.normalize_cells(x)
    got = rc.xnorm.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref), initial=0.0) < 2e-7      # float32 rounding of a float64 result


@pytest.mark.parametrize("n,g,kind", [
    (64, 257, "counts"), (33, 2000, "counts"), (17, 3000, "counts"), (40, 50, "ties"), (9, 1, "counts"),
    (30, 1000, "dense"), (12, 4096, "dense"), (25, 127, "ties"), (26, 128, "counts"), (5, 16000, "counts"),
])
@pytest.mark.parametrize("drop_mode", [0, 1])
def test_rank_normalize_limb_output(eng, n, g, kind, drop_mode):
    """K1 writing the int8 limb planes of the centroid paths directly (mn_rank_normalize_limbs, all three stages,
    dense and CSR input, row / column gathers) = the fp16 planes put through mn_limb_split_cells, bit for bit."""
    from scipy import sparse
    rng = np.random.default_rng(n * 7919 + g + drop_mode)
    x = _rand_matrix(rng, n, g, kind)
    x[0] = 0.0
    if n > 3:
        x[3] = 7.0
    if kind == "dense":
        x[:, ::5] = 0.0
    rows = eng._i32(rng.permutation(n).astype(np.int32))
    Xd = eng.to_device(x)
    ref = eng.rank_normalize(Xd, row_index=rows, drop_mode=drop_mode)
    a1r, a0r, ldk, inv64r = eng.cell_limbs(ref)
    for src in (Xd, eng.to_device_matrix(sparse.csr_matrix(x))):
        got = eng.rank_normalize(src, row_index=rows, drop_mode=drop_mode, limbs=True)
        a1, a0, ldk2, inv64 = eng.cell_limbs(got)
        torch.cuda.synchronize()
        assert got.d_hi is None and ldk2 == ldk
        assert torch.equal(a1, a1r) and torch.equal(a0, a0r), "limb planes differ"
        assert torch.equal(inv64, inv64r) and torch.equal(got.inv_norm, ref.inv_norm)
        assert torch.equal(got.flags & 7, ref.flags & 7)
    # the planes reassemble to the exact centred ranks: d = 128 a1 + a0
    d = (1 if g <= 127 else 128) * a1r.cpu().numpy().astype(np.int64) + a0r.cpu().numpy().astype(np.int64)
    keep = (ref.flags.cpu().numpy() & 4) == 0
    want = O.rank2_rows(x[rows.cpu().numpy()]) - (g + 1)
    assert np.array_equal(d[keep][:, :g], want[keep]) and not d[:, g:].any() and not d[~keep].any()
    if g > 20:
        cols = eng._i32(np.sort(rng.choice(g, size=g // 2, replace=False)).astype(np.int32))
        ref = eng.rank_normalize(Xd, row_index=rows, col_index=cols, drop_mode=drop_mode)
        got = eng.rank_normalize(Xd, row_index=rows, col_index=cols, drop_mode=drop_mode, limbs=True)
        for u, v in zip(eng.cell_limbs(ref)[:2] + eng.cell_limbs(ref)[3:], eng.cell_limbs(got)[:2] + eng.cell_limbs(got)[3:]):
            assert torch.equal(u, v)


def test_centroids_from_limbs_equal_planes(eng):
    """K2 on the limb planes (mn_centroids_fixed_i8) = K2 on the fp16 planes: the same integer sums."""
    rng = np.random.default_rng(77)
    for n, g, L in [(500, 300, 7), (333, 2500, 40), (64, 100, 3)]:
        x = _rand_matrix(rng, n, g, "counts")
        x[5] = 0
        cuts = np.sort(rng.choice(np.arange(1, n), size=L - 1, replace=False))
        row_off = eng._i32(np.concatenate([[0], cuts, [n]]).astype(np.int32))
        Xd = eng.to_device(x)
        a = eng.centroids_fixed(eng.rank_normalize(Xd), row_off, L, extra_rows=1)
        b = eng.centroids_fixed(eng.rank_normalize(Xd, limbs=True), row_off, L, extra_rows=1)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_rank_normalize_golden_and_gather(eng):
    spec, inp, ref = C.load_case("normalize")
    x = inp["X"]
    rc = eng.rank_normalize(eng.to_device(x), want_xnorm=True, want_rank2=True)
    assert np.array_equal(rc.rank2.cpu().numpy(), ref["rank2"])
    np.testing.assert_allclose(rc.xnorm.cpu().numpy(), ref["xnorm"], atol=2e-7, rtol=0)
    # row / column gathers
    rng = np.random.default_rng(5)
    rows = rng.permutation(x.shape[0])[:40].astype(np.int32)
    cols = np.sort(rng.choice(x.shape[1], size=91, replace=False)).astype(np.int32)
    rc2 = eng.rank_normalize(eng.to_device(x), row_index=eng._i32(rows), col_index=eng._i32(cols), want_rank2=True)
    assert np.array_equal(rc2.rank2.cpu().numpy(), O.rank2_rows(x[np.ix_(rows, cols)]))


def test_drop_mode_row_sum(eng):
    x = np.array([[0, 0, 0, 0], [1, 2, 0, 0], [-3, 1, 1, 0], [2, 2, 2, 2]], dtype=np.float32)
    rc = eng.rank_normalize(eng.to_device(x), drop_mode=1)
    f = rc.flags.cpu().numpy()
    assert list(f & 1) == [1, 0, 0, 1]          # zero variance
    assert list((f >> 1) & 1) == [1, 0, 1, 0]   # sum <= 0
    assert list((f >> 2) & 1) == [1, 0, 1, 1]   # dropped


def _planes(rc, g):
    hi = rc.d_hi.cpu().numpy().astype(np.float64)[:, :g]
    if rc.d_lo is not None:
        hi = hi + rc.d_lo.cpu().numpy().astype(np.float64)[:, :g]
    return hi


@pytest.mark.parametrize("n,g,L", [(500, 130, 7), (300, 2500, 3)])
def test_centroids_and_group_sum(eng, n, g, L):
    rng = np.random.default_rng(1)
    x = _rand_matrix(rng, n, g, "counts")
    x[5] = 0
    rc = eng.rank_normalize(eng.to_device(x))
    cuts = np.sort(rng.choice(np.arange(1, n), size=L - 1, replace=False))
    row_off = np.concatenate([[0], cuts, [n]]).astype(np.int32)
    Cd, nv = eng.centroids(rc, eng._i32(row_off), L, extra_rows=2)
    group = (np.arange(L) % 2).astype(np.int32)
    eng.group_sum_into(Cd, nv, L, eng._i32(group), 2, g)
    xn = _planes(rc, g) * rc.inv_norm.cpu().numpy().astype(np.float64)[:, None]
    ref = np.stack([xn[row_off[l]:row_off[l + 1]].sum(axis=0) for l in range(L)])
    got = Cd.cpu().numpy()[:, :g]
    np.testing.assert_allclose(got[:L], ref, atol=2e-6, rtol=1e-6)
    kept = (rc.inv_norm.cpu().numpy() > 0)
    assert np.array_equal(nv.cpu().numpy()[:L], [kept[row_off[l]:row_off[l + 1]].sum() for l in range(L)])
    for q in range(2):
        np.testing.assert_allclose(got[L + q], got[:L][group == q].sum(axis=0), atol=1e-5, rtol=1e-6)
        assert nv.cpu().numpy()[L + q] == nv.cpu().numpy()[:L][group == q].sum()


@pytest.mark.parametrize("m,nn,g", [(300, 42, 120), (1000, 200, 2500), (129, 257, 64), (5000, 13, 2000),
                                    (4097, 128, 333)])
def test_votes_fast_tcgen05_vs_reference(eng, m, nn, g):
    """tcgen05 GEMM vs float64 NumPy on the very same fp16 planes."""
    rng = np.random.default_rng(m + nn + g)
    x = _rand_matrix(rng, m, g, "counts")
    rc = eng.rank_normalize(eng.to_device(x))
    cent = rng.normal(size=(nn, g)).astype(np.float32) * rng.uniform(0.01, 50, size=(nn, 1)).astype(np.float32)
    Cd = torch.zeros((nn, rc.ldd), dtype=torch.float32, device="cuda")
    Cd[:, :g] = eng.to_device(cent)
    b_hi, b_lo, inv_scale = eng.split_f16(Cd, g)
    addend = eng.to_device(rng.uniform(0, 100, size=nn).astype(np.float32))
    out = eng.votes_fast(rc, b_hi, b_lo, inv_scale, addend, nn)
    torch.cuda.synchronize()
    a = _planes(rc, g)
    b = (b_hi.cpu().numpy().astype(np.float64) + b_lo.cpu().numpy().astype(np.float64))[:, :g]
    ref = (a @ b.T) * rc.inv_norm.cpu().numpy().astype(np.float64)[:, None] * inv_scale.cpu().numpy().astype(np.float64)[None, :]
    ref = ref + addend.cpu().numpy().astype(np.float64)[None, :]
    got = out.cpu().numpy().T.astype(np.float64)
    scale = np.abs(ref - addend.cpu().numpy()[None, :]).max() + 1e-30
    err = np.abs(got - ref).max() / scale
    # FP32 accumulation of K ~ 2-3k exact products: IEEE sequential summation would give ~sqrt(K)*2^-24 = 3e-6
    assert err < 4e-6, f"tcgen05 votes differ: rel err {err:.3e}"
    # the split itself reproduces the float32 centroids to ~2^-22
    np.testing.assert_allclose(b * inv_scale.cpu().numpy().astype(np.float64)[:, None], cent, rtol=0,
                               atol=np.abs(cent).max(axis=1, keepdims=True).max() * 2.0 ** -21)


@pytest.mark.parametrize("m,nn,g", [(300, 42, 120), (1000, 200, 2500), (129, 257, 64), (5000, 13, 2000),
                                    (4097, 128, 333), (700, 1057, 3000), (260, 5, 9000), (900, 30, 127)])
def test_votes_exact_int8_limbs(eng, m, nn, g):
    """The int8 limb form (mn_limb_split_* + mn_votes_fast_i8, the product path of the centroid votes): the limb
    planes reproduce the integers they were split from, and the votes equal the float64 product of the integer
    ranks with the quantised centroid to float32 rounding -- the integer dot products are exact."""
    rng = np.random.default_rng(m + nn + g)
    x = _rand_matrix(rng, m, g, "counts")
    x[3] = 0                                                     # a dropped cell
    rc = eng.rank_normalize(eng.to_device(x))
    cent = rng.normal(size=(nn, g)).astype(np.float32) * rng.uniform(0.01, 50, size=(nn, 1)).astype(np.float32)
    Cd = torch.zeros((nn + 1, rc.ldd), dtype=torch.float32, device="cuda")
    Cd[:nn, :g] = eng.to_device(cent)
    addend = eng.to_device(rng.uniform(0, 100, size=nn).astype(np.float32))
    out = eng.votes_exact(rc, Cd, addend, nn)
    torch.cuda.synchronize()
    a1, a0, ldk, inv64 = eng.cell_limbs(rc)
    d = _planes(rc, g).astype(np.int64)
    lim = (128 if g > 127 else 1) * a1.cpu().numpy().astype(np.int64) + a0.cpu().numpy().astype(np.int64)
    assert np.array_equal(lim[:, :g], d) and not lim[:, g:].any(), "limb planes must reproduce the integer ranks"
    assert a0.cpu().numpy().min() >= -64 and a0.cpu().numpy().max() <= 63
    ss = (d.astype(np.float64) ** 2).sum(axis=1)
    inv_ref = np.where(ss > 0, 1.0 / np.sqrt(np.maximum(ss, 1.0)), 0.0)
    assert np.array_equal(inv64.cpu().numpy(), inv_ref), "double-precision scale from the exact integer norm"
    # the centroid as the kernel quantises it: 28 bits of the row's largest entry, in balanced base-128 digits;
    # the weight-1 cross term of the lowest digits of both operands is the one product that is not formed
    e = np.frexp(np.abs(cent).max(axis=1))[1]
    sc = np.ldexp(1.0, 27 - e)[:, None]
    bq = np.rint(cent.astype(np.float64) * sc).astype(np.int64)
    a0_ = ((d + 64) % 128) - 64 if g > 127 else np.zeros_like(d)
    b0_ = ((bq + 64) % 128) - 64
    dot = d @ bq.T - a0_ @ b0_.T
    ref = dot.astype(np.float64) * inv_ref[:, None] / sc.T + addend.cpu().numpy().astype(np.float64)[None, :]
    got = out.cpu().numpy().T
    assert np.array_equal(got, ref.astype(np.float32)), \
        f"exact path off by {np.abs(got - ref).max():.3e} (values up to {np.abs(ref).max():.3e})"
    assert np.abs(bq / sc - cent).max() <= np.abs(cent).max() * 2.0 ** -27
    assert np.abs(a0_ @ b0_.T).max() <= 6.0 * 4096.0 * np.sqrt(g), "the digit product that is not formed is a zero-mean sum"
    # node degree in double (mn_votes_den64) and the centred ratio of the epilogue
    S = min(nn, 3)
    den = eng.votes_den64(rc, Cd[:S], addend[:S].contiguous(), S, cuda_cores=True)
    den_ref = (d @ bq[:S].T).astype(np.float64) * inv_ref[:, None] / sc.T[:, :S] + addend.cpu().numpy().astype(np.float64)[None, :S]
    torch.cuda.synchronize()
    assert np.array_equal(den.cpu().numpy().T, den_ref), "node degree: exact integer dot products, scaled in double"
    # the product path forms the same eight digit products on the int8 tensor cores (two launches, the second adds
    # the weight-1 product into the first's doubles): equal to the last bit or two of a double
    den_tc = eng.votes_den64(rc, Cd[:S], addend[:S].contiguous(), S)
    torch.cuda.synchronize()
    np.testing.assert_allclose(den_tc.cpu().numpy().T, den_ref, rtol=1e-15, atol=np.abs(den_ref).max() * 1e-15)
    col_den = eng._i32((np.arange(nn) % S).astype(np.int32))
    scale = eng.to_device(rng.uniform(0.5, 2.0, size=nn))
    rat = eng.votes_exact(rc, Cd, addend, nn, den=den, col_den=col_den, col_scale=scale)
    torch.cuda.synchronize()
    want = ref / den.cpu().numpy().T[:, np.arange(nn) % S] * scale.cpu().numpy()[None, :] - 1.0
    np.testing.assert_allclose(rat.cpu().numpy().T, want.astype(np.float32), rtol=2e-7, atol=2e-7)


@pytest.mark.parametrize("n_seg,seg_n,ncols,n_labels,ties", [
    (1, 1000, 5, 4, False), (3, 2500, 9, 12, True), (2, 9000, 3, 6, False), (1, 37, 2, 2, True), (2, 5000, 4, 300, True),
])
def test_auroc_kernel_matches_oracle(eng, n_seg, seg_n, ncols, n_labels, ties):
    rng = np.random.default_rng(seg_n + ncols)
    M = n_seg * seg_n
    votes = rng.normal(size=(ncols, M)).astype(np.float32)
    if ties:
        votes = np.round(votes, 1)
        votes[0, : seg_n // 2] = 0.25                      # one giant tie run
        if ncols > 1:
            votes[1, :] = 1.0                              # a fully tied column
    denom = rng.uniform(0.5, 2.0, size=(2, M)).astype(np.float32)
    col_denom = (np.arange(ncols) % 2).astype(np.int32)
    col_denom[-1] = -1
    per_seg = n_labels // n_seg
    lab = np.concatenate([np.sort(rng.integers(0, per_seg, size=seg_n)) + s * per_seg for s in range(n_seg)]).astype(np.int32)
    excl = rng.random(M) < 0.05
    lab_dev = np.where(excl, -1, lab).astype(np.int32)
    seg_off = (np.arange(n_seg + 1) * seg_n).astype(np.int32)
    auc, n_pos = eng.auroc(eng.to_device(votes), ncols, eng.to_device(denom), eng._i32(col_denom), eng._i32(seg_off),
                           n_seg, eng._i32(lab_dev), n_labels)
    torch.cuda.synchronize()
    got = auc.cpu().numpy()
    ref = np.full((n_labels, ncols), np.nan)
    for s in range(n_seg):
        sl = slice(s * seg_n, (s + 1) * seg_n)
        keep = ~excl[sl]
        key = votes[:, sl].copy()
        for c in range(ncols):
            if col_denom[c] >= 0:
                key[c] = key[c] / denom[col_denom[c], sl]              # float32 division, as on the device
        present, codes = O.encode(lab[sl][keep])
        roc, _, _ = O.compute_aurocs(key[:, keep].T.astype(np.float64), codes, len(present))
        ref[np.asarray(present, dtype=int)] = roc
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref), initial=0) == 0.0, "rank sums are integers: AUROC must be bit-identical"
    cnt = np.bincount(lab[~excl], minlength=n_labels)
    assert np.array_equal(n_pos.cpu().numpy(), cnt)


def _exact_auroc(key_f32, lab, n_labels):
    """Rank-sum AUROC of every label from the exact average-tie ranks (integers x 2), utils.py:172-178."""
    keep = lab >= 0
    k, l = key_f32[keep], lab[keep]
    r = __import__("scipy.stats", fromlist=["rankdata"]).rankdata(k.astype(np.float64), method="average")
    out = np.full(n_labels, np.nan)
    n = k.shape[0]
    for c in np.unique(l):
        pos = l == c
        n_p = float(pos.sum())
        roc = (r[pos].sum() / n_p - (n_p + 1.0) / 2.0) / float(n - n_p) if n - n_p > 0 else np.nan
        out[c] = roc
    return out


@pytest.mark.parametrize("shape", ["skewed_71k", "giant_tie_60k", "many_tie_blocks", "wide_range", "two_values",
                                   "dense_cluster_100k", "unsorted_labels", "million"])
def test_auroc_counting_kernel_adversarial(eng, shape):
    """The counting kernel (auroc_count.cu) at the sizes and distributions that stress its binning: several rounds
    per segment, a long-tailed vote distribution, a tie block larger than one round, thousands of tie blocks,
    keys spread over many binades and both signs, one coarse bin holding more than a round of distinct keys
    (recursive refinement), labels in arbitrary cell order, a 1M-cell segment.  Bit-identical to exact ranks."""
    rng = np.random.default_rng(__import__("zlib").crc32(shape.encode()))
    ncols, n_labels = 3, 37
    if shape == "skewed_71k":
        n = 71429
        votes = (0.0063 + rng.gamma(1.2, 0.0002, size=(ncols, n))).astype(np.float32)
        votes[:, ::150] += 0.003
    elif shape == "giant_tie_60k":
        n = 90000
        votes = rng.normal(size=(ncols, n)).astype(np.float32)
        votes[:, rng.permutation(n)[:60000]] = 0.125
    elif shape == "many_tie_blocks":
        n = 50000
        votes = np.round(rng.normal(size=(ncols, n)), 2).astype(np.float32)
        votes[1] = np.round(votes[1], 0)
    elif shape == "wide_range":
        n = 66000
        votes = (rng.normal(size=(ncols, n)) * np.exp(rng.uniform(-30, 30, size=(ncols, n)))).astype(np.float32)
        votes[2, :100] = 0.0
        votes[2, 100:200] = -0.0
    elif shape == "two_values":
        n = 120000
        votes = rng.integers(0, 2, size=(ncols, n)).astype(np.float32)
    elif shape == "dense_cluster_100k":
        n = 150000                        # 100k distinct keys inside a sliver of the range: one coarse bin > one round
        votes = rng.uniform(0.0, 1000.0, size=(ncols, n)).astype(np.float32)
        votes[:, :100000] = (500.0 + rng.uniform(0.0, 0.01, size=(ncols, 100000))).astype(np.float32)
    elif shape == "unsorted_labels":
        n = 30000
        votes = rng.normal(size=(ncols, n)).astype(np.float32)
    else:
        n = 1000000
        votes = rng.normal(size=(ncols, n)).astype(np.float32)
    lab = np.sort(rng.integers(0, n_labels, size=n)).astype(np.int32)
    if shape == "unsorted_labels":
        lab = rng.integers(0, n_labels, size=n).astype(np.int32)
    lab[rng.random(n) < 0.03] = -1
    seg_off = np.array([0, n], np.int32)
    auc, n_pos = eng.auroc(eng.to_device(votes), ncols, None, None, eng._i32(seg_off), 1, eng._i32(lab), n_labels)
    torch.cuda.synchronize()
    got = auc.cpu().numpy()
    for c in range(ncols):
        ref = _exact_auroc(votes[c], lab, n_labels)
        assert np.array_equal(np.isnan(got[:, c]), np.isnan(ref)), (shape, c)
        assert np.nanmax(np.abs(got[:, c] - ref), initial=0) == 0.0, (shape, c)
    assert np.array_equal(n_pos.cpu().numpy(), np.bincount(lab[lab >= 0], minlength=n_labels))


def test_auroc_counting_equals_sort_kernel(eng, monkeypatch):
    """Counting kernel == radix-sort kernel (the form kept for > 2,048 labels), many segments of uneven length,
    node-degree denominators, 300 labels."""
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 30000, size=6)
    lens[2] = 0
    M, ncols, n_labels = int(lens.sum()), 7, 300
    seg_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    votes = (rng.normal(size=(ncols + 2, M)) * 0.01 + 1.0).astype(np.float32)
    lab = np.concatenate([np.sort(rng.integers(0, 50, size=k)) + 50 * s for s, k in enumerate(lens)]).astype(np.int32)
    lab[rng.random(M) < 0.02] = -1
    col_denom = eng._i32((np.arange(ncols) % 2).astype(np.int32))
    vd = eng.to_device(votes)
    a1, p1 = eng.auroc(vd, ncols, vd[ncols:], col_denom, eng._i32(seg_off), len(lens), eng._i32(lab), n_labels)
    monkeypatch.setenv("MN_AUROC_SORT", "1")
    a2, p2 = eng.auroc(vd, ncols, vd[ncols:], col_denom, eng._i32(seg_off), len(lens), eng._i32(lab), n_labels)
    monkeypatch.delenv("MN_AUROC_SORT")
    torch.cuda.synchronize()
    assert torch.equal(p1, p2)
    assert torch.equal(torch.isnan(a1), torch.isnan(a2))
    assert torch.equal(torch.nan_to_num(a1), torch.nan_to_num(a2))


def test_one_vs_best_kernel_matches_oracle(eng):
    rng = np.random.default_rng(3)
    n_seg, per_seg, ncols = 2, 7, 11
    sizes = rng.integers(20, 120, size=n_seg * per_seg)
    lab = np.repeat(np.arange(n_seg * per_seg), sizes).astype(np.int32)
    M = lab.shape[0]
    centers = rng.normal(size=(ncols, n_seg * per_seg))
    votes = (centers[:, lab] + rng.normal(size=(ncols, M))).astype(np.float32)
    votes = np.round(votes, 2)
    lab_row_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    seg_lab_off = (np.arange(n_seg + 1) * per_seg).astype(np.int32)
    seg_off = lab_row_off[seg_lab_off]
    vd = eng.to_device(votes)
    auc, _ = eng.auroc(vd, ncols, None, None, eng._i32(seg_off), n_seg, eng._i32(lab), n_seg * per_seg)
    out = eng.one_vs_best(vd, ncols, None, None, eng._i32(seg_lab_off), n_seg, eng._i32(lab_row_off), eng._i32(lab),
                          n_seg * per_seg, auc)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for s in range(n_seg):
        sl = slice(seg_off[s], seg_off[s + 1])
        codes = lab[sl] - s * per_seg
        roc, _, _ = O.compute_aurocs(votes[:, sl].T.astype(np.float64), codes, per_seg)
        ref = O.one_vs_best(votes[:, sl].T.astype(np.float64), codes, roc)
        blk = got[s * per_seg:(s + 1) * per_seg]
        assert np.array_equal(np.isnan(blk), np.isnan(ref))
        assert np.nanmax(np.abs(blk - ref)) == 0.0


@pytest.mark.parametrize("n,g", [(700, 96), (1500, 300), (640, 2100)])
def test_corr_hist_and_votes(eng, n, g):
    """Pass 1 counts every pair i<j exactly once into the right bin; pass 2's integer vote
    sums equal the LUT-mapped network computed on the host from the same rho values."""
    rng = np.random.default_rng(n + g)
    x, s, c = __import__("pymn_b200.synth", fromlist=["make_counts"]).make_counts(n + g, 2, n // 2, g, 4)
    x[7] = 0                                                # a dropped cell
    rc = eng.rank_normalize(eng.to_device(x))
    L = 5
    label = rng.integers(0, L, size=n).astype(np.int32)
    nb = 4096
    votes, degree, hist, lut = eng.corr_network_votes(rc, eng._i32(label), L, nbins=nb)
    torch.cuda.synchronize()
    d = _planes(rc, g)
    inv = rc.inv_norm.cpu().numpy()
    kept = inv > 0
    # the kernels' bin positions, restated on the host (tests/lut_emulation.py follows rho_q / w_from_q)
    q = E.q_matrix(d, inv, nb)
    ref_hist = E.hist_upper(q, kept, nb)
    got_hist = hist.cpu().numpy()
    m = int(kept.sum())
    assert got_hist.sum() == m * (m - 1) // 2, "every kept pair exactly once"
    # FP32 accumulation in the tensor core may move a value across a bin edge into the
    # neighbouring bin: the cumulative counts must agree up to a few such crossings
    assert np.abs(np.cumsum(got_hist - ref_hist)).max() <= max(4, 2e-3 * ref_hist.sum())
    if g * g * g < (1 << 24):
        assert np.array_equal(got_hist, ref_hist), "all partial sums are exact integers < 2^24: histogram must be exact"
    lut_h = lut.cpu().numpy().astype(np.int64)
    assert np.array_equal(lut_h, E.build_lut(got_hist, n)), "packed CDF table differs from its host restatement"
    w0 = lut_h[:nb + 1] >> 11
    assert np.all(np.diff(w0) >= 0) and w0[0] == 0 and w0[-1] <= (1 << 20)
    # the kernels' weight map: 16.16 fixed-point bin position -> integer interpolation inside the bin
    w = np.triu(E.w_from_lut(lut_h, q), 1)
    w = w + w.T
    w[~kept, :] = 0
    w[:, ~kept] = 0
    np.fill_diagonal(w, 0)
    ref_deg = w.sum(axis=1)
    ref_votes = np.stack([w[:, label == l].sum(axis=1) for l in range(L)], axis=1)
    gd = degree.cpu().numpy()
    gv = votes.cpu().numpy()
    assert np.array_equal(gv.sum(axis=1), gd), "votes over all labels add up to the node degree (exact integers)"
    # g <= 256: every partial sum of the tensor core is an exact integer and so are the sums below.  Beyond, the
    # FP32 accumulator truncates (one chain of ceil(g / 16) adds per plane product: ~6e-6 relative at g = 2,000,
    # four plane products in that one chain at g = 2,100: ~5e-5), which moves rho by that much against the exact integer dot products of
    # the host restatement -- the effect on the AUROC table is what tests/test_gpu_headline.py bounds
    tol_w = 1e-5 if g <= 2048 else 1e-4
    rel = np.abs(gd - ref_deg).max() / ref_deg.max()
    assert rel < tol_w, f"node degree off by {rel:.2e}"
    rel_v = np.abs(gv - ref_votes).max() / ref_votes.max()
    assert rel_v < tol_w, f"votes off by {rel_v:.2e}"
    # the rho spill (pass 1 keeps the bin positions in HBM, pass 2 streams them back) is only a
    # different route for the same integers: none / partial / complete spill give identical sums
    for tiles in (0, 3):
        v2, d2, h2, _ = eng.corr_network_votes(rc, eng._i32(label), L, nbins=nb, spill_tiles=tiles)
        torch.cuda.synchronize()
        assert torch.equal(h2, hist) and torch.equal(v2, votes) and torch.equal(d2, degree), f"spill_tiles={tiles}"


def test_corr_gemm_forms_identical(eng, monkeypatch):
    """Single CTAs, CTA pairs (default) and quad clusters (two pairs, B tile TMA-multicast) are three
    schedules of the same arithmetic: histogram, LUT and vote sums must be bit-identical -- with and
    without the rho spill.  Also the 128-column tile form selected by nbins > 16384."""
    make_counts = __import__("pymn_b200.synth", fromlist=["make_counts"]).make_counts
    x, s, c = make_counts(99, 2, 900, 200, 4)
    x[11] = 0
    rc = eng.rank_normalize(eng.to_device(x))
    rng = np.random.default_rng(5)
    L = 6
    label = np.sort(rng.integers(0, L, size=x.shape[0])).astype(np.int32)
    ref = eng.corr_network_votes(rc, eng._i32(label), L, nbins=8192)
    torch.cuda.synchronize()
    for ctas in ("1", "4"):
        monkeypatch.setenv("MN_GEMM_CTAS", ctas)
        for tiles in (None, 0, 5):
            got = eng.corr_network_votes(rc, eng._i32(label), L, nbins=8192, spill_tiles=tiles)
            torch.cuda.synchronize()
            for a, b, what in zip(got, ref, ("votes", "degree", "hist", "lut")):
                assert torch.equal(a, b), f"ctas={ctas} spill_tiles={tiles}: {what} differs"
    monkeypatch.delenv("MN_GEMM_CTAS")
    # 32768 bins: 128 x 128 tiles, single CTAs; self-consistency across the spill settings
    big = eng.corr_network_votes(rc, eng._i32(label), L, nbins=32768)
    torch.cuda.synchronize()
    m = int((rc.inv_norm > 0).sum().item())
    assert int(big[2].sum().item()) == m * (m - 1) // 2
    assert torch.equal(big[0].sum(dim=1), big[1])
    for tiles in (0, 3):
        got = eng.corr_network_votes(rc, eng._i32(label), L, nbins=32768, spill_tiles=tiles)
        torch.cuda.synchronize()
        assert torch.equal(got[0], big[0]) and torch.equal(got[1], big[1]) and torch.equal(got[2], big[2])


def test_corr_votes_heavy_ties_mid_rank(eng):
    """Six genes: rho takes a handful of values, single bins hold far more than 0.2 % of all pairs; such a
    bin is a tie block and gets its mid-rank weight (rise 0), as bottleneck.rankdata ranks ties (utils.py:45).
    Histogram, table and vote sums equal the host restatement (tests/lut_emulation.py) and do not depend on
    the spill.  (How close that scheme is to the reference's exact ranking: tests/test_lut_scheme.py.)"""
    rng = np.random.default_rng(17)
    n, g, L, nb = 900, 6, 4, 4096
    x = rng.poisson(1.0, size=(n, g)).astype(np.float32)
    rc = eng.rank_normalize(eng.to_device(x))
    label = np.sort(rng.integers(0, L, size=n)).astype(np.int32)
    votes, degree, hist, lut = eng.corr_network_votes(rc, eng._i32(label), L, nbins=nb)
    torch.cuda.synchronize()
    lut_h = lut.cpu().numpy().astype(np.int64)
    assert lut_h[nb + 1] >= 3, "this case is meant to exercise tie-block bins (rise 0, mid-rank weight)"
    d = _planes(rc, g)
    inv = rc.inv_norm.cpu().numpy()
    kept = inv > 0
    q = E.q_matrix(d, inv, nb)
    # six genes: every dot product is an exact integer; only a last-bit difference of the fmaf emulated in
    # float64 could move a pair across a bin edge
    assert np.abs(np.cumsum(hist.cpu().numpy() - E.hist_upper(q, kept, nb))).max() <= 2
    assert np.array_equal(lut_h, E.build_lut(hist.cpu().numpy(), n))
    w = np.triu(E.w_from_lut(lut_h, q), 1)
    w = w + w.T
    w[~kept, :] = 0
    w[:, ~kept] = 0
    np.fill_diagonal(w, 0)
    # (equal up to a last-bit difference of an fmaf emulated in float64: one unit of q moves a weight by rise / 65536)
    ref_deg = w.sum(axis=1)
    ref_votes = np.stack([w[:, label == l].sum(axis=1) for l in range(L)], axis=1)
    assert np.abs(degree.cpu().numpy() - ref_deg).max() <= 1e-6 * ref_deg.max()
    assert np.abs(votes.cpu().numpy() - ref_votes).max() <= 1e-6 * ref_votes.max()
    assert torch.equal(votes.sum(dim=1), degree)
    v0, d0, h0, _ = eng.corr_network_votes(rc, eng._i32(label), L, nbins=nb, spill_tiles=0)
    torch.cuda.synchronize()
    assert torch.equal(v0, votes) and torch.equal(d0, degree) and torch.equal(h0, hist)


def test_node_degree_gemm_repeatable_at_scale(eng):
    """mn_votes_den_i8 at config-4 scale (500k cells, 3,000 genes, 7 study rows): two runs are bit-identical and
    equal the CUDA-core kernel to the last bits of a double (a tile variant once passed every small test and
    failed here)."""
    from pymn_b200.synth import make_counts_torch
    S = 7
    Xd, _, _ = make_counts_torch(4, S, 71429, 3000, 150, torch.device("cuda", 0))
    rc = eng.rank_normalize(Xd, limbs=True)
    del Xd
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    C = torch.randn((S, rc.ldd), device="cuda", generator=g) * 3.0
    C[:, 3000:] = 0
    add = torch.rand((S,), device="cuda", generator=g) * 100
    ref = eng.votes_den64(rc, C, add, S, cuda_cores=True)
    a = eng.votes_den64(rc, C, add, S)
    b = eng.votes_den64(rc, C, add, S)
    torch.cuda.synchronize()
    assert torch.equal(a, b), "node-degree GEMM is not repeatable"
    # (a few ulps of a double: the two launches round val = dot * scale + addend in two pieces)
    err = ((a - ref).abs() / ref.abs().clamp(min=1.0)).max().item()
    assert err < 1e-12, f"node-degree GEMM differs from the CUDA-core kernel: {err:.2e}"
