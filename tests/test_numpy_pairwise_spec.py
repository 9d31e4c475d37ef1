"""The summation order the CUDA area kernel implements (oracle/pairwise.py) IS numpy's: bit-for-bit vs np.sum /
np.nanmean on the array shapes the reference produces (1-D lists of correlations)."""
import numpy as np

from oracle.pairwise import pairwise_sum


def test_pairwise_equals_np_sum_all_lengths():
    rng = np.random.default_rng(0)
    for n in list(range(0, 300)) + [511, 512, 513, 1000, 1023, 1024, 1025, 2049, 4097, 6561]:
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
        assert pairwise_sum(list(a)) == (np.sum(a) if n else 0.0), n


def test_nanmean_of_list_is_pairwise_over_whole_array():
    rng = np.random.default_rng(1)
    for n in (5, 8, 9, 63, 64, 127, 128, 129, 200, 277, 700):
        a = rng.uniform(-1, 1, n)
        assert np.nanmean(list(a)) == pairwise_sum(list(a)) / n
        b = a.copy()
        b[rng.integers(0, n, max(1, n // 9))] = np.nan          # NaN -> 0, divide by the non-NaN count
        z = np.where(np.isnan(b), 0.0, b)
        assert np.nanmean(list(b)) == pairwise_sum(list(z)) / np.sum(~np.isnan(b))


def test_row_reduction_of_2d_equals_1d():
    """oracle/network.py sums rows of a gathered 2-D block one 1-D row at a time, like the reference's lists."""
    rng = np.random.default_rng(2)
    G = rng.uniform(-1, 1, (17, 301))
    for q in range(G.shape[0]):
        assert np.nanmean(G[q]) == pairwise_sum(list(G[q])) / G.shape[1]
