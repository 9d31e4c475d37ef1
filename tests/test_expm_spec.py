"""The expm algorithm the GP kernel implements (oracle/expm_spec.py) vs scipy.linalg.expm, on the graph-Laplacian
matrices the forecaster builds (M = |cov|, zero row sums) over the whole (l, scale) range of the scripts."""
import numpy as np
from scipy.linalg import expm

from oracle.expm_spec import expm_spec, norm1


def laplacian(rng, Np, n, scale):
    Xd = rng.standard_normal((n, Np)) * scale
    M = np.atleast_2d(np.abs(np.cov(Xd, rowvar=False, bias=True)))
    np.fill_diagonal(M, 0)
    np.fill_diagonal(M, -M.sum(axis=0))
    return M


def test_matches_scipy_rounding_level_when_well_conditioned():
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(200):
        Np, n = int(rng.integers(2, 80)), int(rng.integers(6, 42))
        A = 10 ** rng.uniform(-9, 3) * laplacian(rng, Np, n, 10 ** rng.uniform(-1, 2))
        E = expm(A)
        F, m, s = expm_spec(A, info=True)
        seen.add(m)
        err = np.abs(E - F).max() / np.abs(E).max()
        assert err <= max(1e-13, 64 * 2.0 ** s * 2.0 ** -53), (trial, m, s, err)
    assert seen == {3, 5, 7, 9, 13}


def test_large_norm_deviation_is_scipys_own_sensitivity():
    """||lM|| up to 1e13 (July's l = 3.125433e+10, north/July1st.py:169): scipy's result moves by ~2^s*eps under a
    1-ulp perturbation of its input; the restated algorithm stays within a small multiple of that, i.e. it picks
    the same (m, s) and differs only by amplified rounding."""
    rng = np.random.default_rng(1)
    for trial in range(12):
        Np, n = int(rng.integers(20, 80)), int(rng.integers(6, 42))
        A = 3.125433e+10 * 10 ** rng.uniform(-6, -1) * laplacian(rng, Np, n, 10 ** rng.uniform(0, 2))
        E = expm(A)
        F, m, s = expm_spec(A, info=True)
        P = A * (1 + (rng.integers(0, 2, A.shape) * 2 - 1) * 2.0 ** -52)
        P = (P + P.T) / 2
        E2 = expm(P)
        sc = np.abs(E).max()
        own = np.abs(E - E2).max() / sc
        mine = np.abs(E - F).max() / sc
        assert m == 13 and s > 10
        assert mine <= 50 * max(own, 2.0 ** s * 2.0 ** -53), (norm1(A), s, mine, own)
