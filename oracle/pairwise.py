"""ORACLE (test infrastructure) -- numpy's pairwise summation restated in plain Python.

numpy/_core/src/umath/loops_utils.h.src `pairwise_sum_DOUBLE` is the order in which `np.sum` / `np.nanmean`
add a contiguous 1-D float64 array; the reference's growth/merge decisions (ComplexNetworks.py:113, :250, :252)
compare such sums, and the CUDA kernel (csrc/common.cuh: sie_pw_leaf8 / sie_pw_sum8) evaluates them in this
order.  tests/test_numpy_pairwise_spec.py pins this restatement against `np.sum` itself."""


def pairwise_sum(a, lo=0, n=None):
    if n is None:
        n = len(a)
    if n < 8:
        res = 0.0
        for i in range(n):
            res += a[lo + i]
        return res
    if n <= 128:
        r = [a[lo + j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[lo + i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[lo + i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a, lo, n2) + pairwise_sum(a, lo + n2, n - n2)
