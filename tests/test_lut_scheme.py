"""The rank -> weight scheme of the full-network path (histogram of rho + packed CDF table with mid-rank
tie blocks) held against the oracle's exact global ranking (utils.py:35-53) on the CPU.

`tests/lut_emulation.py` restates the kernels' integer arithmetic (the GPU tests check the kernels against
it bit for bit); here that arithmetic is compared with the float64 reference on the inputs where a binned
rank is weakest: small, sparse gene sets whose correlations are heavily tied (a tie block used
to be shifted by up to half its size).

Reference noise: on such data np.corrcoef's float64 rounding splits blocks of mathematically equal
correlations into sub-blocks (ranked in rounding order), so the reference's own AUROCs move by
`noise` = |oracle - oracle with correlations rounded to 1e-11| when the BLAS changes.  The scheme has to stay
within max(4.5 rank flips, 3 x noise) of the oracle.
"""
import numpy as np
import pytest

from oracle import mn_oracle as O
from pymn_b200.synth import make_counts
from tests import lut_emulation as E


@pytest.mark.parametrize("g", [8, 12, 20, 50, 200])
def test_scheme_vs_exact_rank_small_sparse_gene_sets(g):
    x, s, c = make_counts(3, 3, 200, 400, 5)
    rng = np.random.default_rng(g)
    quantum = 1.0 / (40 * 160)
    for rep in range(2):
        sub = x[:, rng.choice(400, size=g, replace=False)]
        ref, _ = E.us_default_from_w(O.create_nw_spearman(sub), s, c)
        merged, _ = E.us_default_from_w(E.tie_merged_reference(sub), s, c)
        got, _ = E.us_default_from_w(E.weights(sub, 16384), s, c)
        noise = float(np.nanmax(np.abs(merged - ref)))
        d = float(np.nanmax(np.abs(got - ref)))
        assert d <= max(4.5 * quantum, 3.0 * noise) + 1e-12, f"g={g} rep={rep}: {d:.2e} (reference noise {noise:.2e})"


def test_tie_block_gets_mid_rank():
    """A bin that holds >= 0.2 % of all pairs is one tie block: weight = the block's mid-rank, no interpolation."""
    hist = np.zeros(1024, np.int64)
    hist[10], hist[500], hist[900] = 50, 500000, 50
    lut = E.build_lut(hist, n_diag=100)
    w0, rise = lut[:1024] >> 11, lut[:1024] & 2047
    total = hist.sum()
    top = 2.0 * total + 100 - 99 * 0.5 - 1
    assert rise[500] == 0 and lut[1025] == 1
    assert w0[500] == int(np.rint((2.0 * 50 + 500000 - 0.5) / top * (1 << 20)))
    assert rise[10] > 0 and rise[900] > 0 and w0[10] == 0
    q = np.array([(500 << 16) + 1, (500 << 16) + 65535])
    assert len(set(E.w_from_lut(lut, q).tolist())) == 1


def test_simple_fractions_are_not_on_bin_edges():
    """rho = k / 2^m used to sit exactly on a bin edge of (rho + 1) * nbins / 2."""
    for nbins in (4096, 16384, 32768):
        for rho in (-0.5, -0.25, 0.25, 0.5, 0.75, 1.0 / 3, 2.0 / 3, -1.0, 1.0):
            pos = rho * float(E.bin_scale(nbins)) + float(E.bin_offset(nbins))
            assert 0.5 < pos < nbins - 0.5
            assert min(pos - np.floor(pos), np.ceil(pos) - pos) > 1e-3, (nbins, rho, pos)
