"""Property tests (hypothesis): host bookkeeping on the CPU, integer kernels on the GPU.

CPU: the cell shard of the multi-GPU centroid paths partitions the cells and reassembles to the global device row
order for every world size.  GPU: K1's doubled ranks equal SciPy's average-tie ranks for arbitrary small rows (both
output forms); K6's AUROCs equal the exact rank-sum definition for arbitrary votes with ties and excluded cells."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from oracle import mn_oracle as O

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@settings(max_examples=60, **COMMON)
@given(n=st.integers(1, 200), world=st.integers(1, 8), n_st=st.integers(1, 4), n_ct=st.integers(1, 5), seed=st.integers(0, 10 ** 6))
def test_cell_shard_partitions_and_reassembles(n, world, n_st, n_ct, seed):
    """Every cell belongs to exactly one rank; a rank's cells are a label-sorted subsequence of the global order;
    gathering the ranks' (padded) local orders through gidx restores the global device row order."""
    from pymn_b200 import engine
    from pymn_b200.utils import label_layout
    rng = np.random.default_rng(seed)
    study = np.array([f"s{v}" for v in rng.integers(0, n_st, size=n)], dtype=object)
    ctype = np.array([f"c{v}" for v in rng.integers(0, n_ct, size=n)], dtype=object)
    lay = label_layout(study, ctype)
    L = len(lay.row_labels)
    shards = [engine.cell_shard(lay, r, world) for r in range(world)]
    owned = np.concatenate([sh.rows for sh in shards])
    assert sorted(owned.tolist()) == list(range(n))
    label_of_input = np.empty(n, np.int64)
    label_of_input[lay.perm] = lay.cell_label
    total = np.zeros(L, np.int64)
    for sh in shards:
        assert len(sh.lab_row_off) == L + 1 and sh.lab_row_off[-1] == len(sh.rows)
        labs = label_of_input[sh.rows]
        assert np.all(np.diff(labs) >= 0)
        total += np.diff(sh.lab_row_off)
        assert sh.max_label_rows == (int(np.diff(sh.lab_row_off).max()) if L else 0)
    assert np.array_equal(total, np.diff(lay.lab_row_off))
    if world > 1:
        per = shards[0].per
        padded = np.full(world * per, -1, np.int64)
        for r, sh in enumerate(shards):
            padded[r * per:r * per + len(sh.rows)] = sh.rows
        for sh in shards:
            assert np.array_equal(padded[sh.gidx], lay.perm)


@pytest.fixture(scope="module")
def eng():
    from pymn_b200 import engine
    return engine


_rows = hnp.arrays(np.float32, st.tuples(st.integers(1, 12), st.integers(1, 300)),
                   elements=st.one_of(st.integers(0, 6).map(float), st.integers(0, 2047).map(float),
                                      st.floats(-50, 50, width=32, allow_nan=False)))


@pytest.mark.gpu
@settings(max_examples=40, **COMMON)
@given(x=_rows)
def test_rank_normalize_property(eng, x):
    """2 * rank of K1 (counting / register-sort / CTA stages, whichever a row takes) equals SciPy's average-tie ranks;
    the int8 limb planes reassemble to the same centred ranks."""
    import torch
    g = x.shape[1]
    Xd = eng.to_device(np.ascontiguousarray(x))
    rc = eng.rank_normalize(Xd, want_rank2=True)
    lim = eng.rank_normalize(Xd, limbs=True)
    torch.cuda.synchronize()
    ref2 = O.rank2_rows(x)
    assert np.array_equal(rc.rank2.cpu().numpy(), ref2)
    a1, a0, _, _ = eng.cell_limbs(lim)
    d = (1 if g <= 127 else 128) * a1.cpu().numpy().astype(np.int64) + a0.cpu().numpy().astype(np.int64)
    keep = (lim.flags.cpu().numpy() & 4) == 0
    assert np.array_equal(d[keep][:, :g], (ref2 - (g + 1))[keep]) and not d[~keep].any() and not d[:, g:].any()


@pytest.mark.gpu
@settings(max_examples=40, **COMMON)
@given(n=st.integers(2, 400), ncols=st.integers(1, 4), n_lab=st.integers(1, 6), levels=st.integers(1, 50),
       seed=st.integers(0, 10 ** 6))
def test_auroc_property(eng, n, ncols, n_lab, levels, seed):
    """K6 against the rank-sum definition (utils.py:172-178) for votes drawn from a few levels (heavy ties) with
    excluded cells: bit-identical (the rank sums are integers)."""
    import torch
    from scipy.stats import rankdata
    rng = np.random.default_rng(seed)
    votes = rng.integers(0, levels, size=(ncols, n)).astype(np.float32) / np.float32(7.0)
    lab = np.sort(rng.integers(0, n_lab, size=n)).astype(np.int32)
    lab_dev = np.where(rng.random(n) < 0.1, -1, lab).astype(np.int32)
    seg = np.array([0, n], dtype=np.int32)
    auc, n_pos = eng.auroc(eng.to_device(votes), ncols, None, None, eng._i32(seg), 1, eng._i32(lab_dev), n_lab)
    torch.cuda.synchronize()
    got = auc.cpu().numpy()
    keep = lab_dev >= 0
    nk = int(keep.sum())
    for c in range(ncols):
        r = rankdata(votes[c][keep].astype(np.float64), method="average") if nk else np.zeros(0)
        for l in range(n_lab):
            pos = lab_dev[keep] == l
            n_p = int(pos.sum())
            if n_p == 0:
                assert np.isnan(got[l, c])
                continue
            with np.errstate(invalid="ignore", divide="ignore"):
                want = (r[pos].sum() / n_p - (n_p + 1) / 2.0) / (nk - n_p)
            assert (np.isnan(want) and np.isnan(got[l, c])) or got[l, c] == want, (l, c, got[l, c], want)
    assert np.array_equal(n_pos.cpu().numpy(), np.bincount(lab_dev[keep], minlength=n_lab))
