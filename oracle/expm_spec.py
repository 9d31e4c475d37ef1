"""ORACLE (test infrastructure) -- the matrix-exponential algorithm the GP kernel implements, in numpy.

scipy.linalg.expm (scipy 1.18.1, compiled `_internal_matfuncs.matrix_exponential`; source not shipped in the
wheel) implements Al-Mohy & Higham (2009), "A New Scaling and Squaring Algorithm for the Matrix Exponential",
SIAM J. Matrix Anal. Appl. 31(3).  This file restates that published algorithm the way csrc/gp.cu evaluates it
(explicit A^2, A^4, A^6 [, A^8, A^10]; exact ||.||_1 of the powers for n < 400; the exact
|| |A|^(2m+1) ||_1 backward-error check by repeated transposed mat-vecs; Pade orders 3,5,7,9,13; 2^-s scaling and
s squarings).  tests/test_expm_spec.py pins it against scipy.linalg.expm: agreement is at rounding level wherever
expm is well conditioned and within scipy's own sensitivity to a 1-ulp input perturbation (~2^s * eps) elsewhere
(SURVEY.md H3).  The reference calls expm at north/June1st.py:264,269."""
import math

import numpy as np
from scipy.linalg import solve

THETA = [1.495585217958292e-002, 2.539398330063230e-001, 9.504178996162932e-001, 2.097847961257068e+000, 4.25]
_U = 2.0 ** -53
COEFF = [_U * 100800., _U * 10059033600., _U * 4487938430976000., _U * 5914384781877411840000.,
         _U * 113250775606021113483283660800000000.]
B = {3: [120., 60., 12., 1.],
     5: [30240., 15120., 3360., 420., 30., 1.],
     7: [17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.],
     9: [17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.],
     13: [64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
          10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.]}


def norm1(A):
    return np.abs(A).sum(axis=0).max()


def absnorm_power(absA, p):
    v = np.ones(absA.shape[0])
    Mt = absA.T
    for _ in range(p):
        v = Mt @ v
    return v.max()


def ell(absA, normA, idx, m):
    t = absnorm_power(absA, 2 * m + 1)
    if not t > 0:
        return 0
    with np.errstate(over="ignore", invalid="ignore"):
        val = math.log2(t / normA / COEFF[idx]) / (2 * m) if np.isfinite(t) else float("inf")
    if not np.isfinite(val):
        return 1 << 20
    return max(int(math.ceil(val)), 0)


def pick(A):
    absA = np.abs(A)
    A2 = A @ A
    A4 = A2 @ A2
    A6 = A4 @ A2
    normA = norm1(A)
    d4 = norm1(A4) ** 0.25
    d6 = norm1(A6) ** (1 / 6.)
    eta0 = max(d4, d6)
    if eta0 < THETA[0] and ell(absA, normA, 0, 3) == 0:
        return 3, 0, (A2, A4, A6, None)
    if eta0 < THETA[1] and ell(absA, normA, 1, 5) == 0:
        return 5, 0, (A2, A4, A6, None)
    A8 = A4 @ A4
    d8 = norm1(A8) ** 0.125
    eta2 = max(d6, d8)
    if eta2 < THETA[2] and ell(absA, normA, 2, 7) == 0:
        return 7, 0, (A2, A4, A6, A8)
    if eta2 < THETA[3] and ell(absA, normA, 3, 9) == 0:
        return 9, 0, (A2, A4, A6, A8)
    d10 = norm1(A4 @ A6) ** 0.1
    eta3 = max(d8, d10)
    eta4 = min(eta2, eta3)
    s = max(int(math.ceil(math.log2(eta4 / THETA[4]))), 0)
    sc = 2.0 ** -s
    s += ell(absA * sc, normA * sc, 4, 13)
    return 13, s, (A2, A4, A6, A8)


def expm_spec(A, info=False):
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    eye = np.eye(n)
    m, s, (A2, A4, A6, A8) = pick(A)
    b = B[m]
    if m == 3:
        U = A @ (b[3] * A2 + b[1] * eye)
        V = b[2] * A2 + b[0] * eye
    elif m == 5:
        U = A @ (b[5] * A4 + b[3] * A2 + b[1] * eye)
        V = b[4] * A4 + b[2] * A2 + b[0] * eye
    elif m == 7:
        U = A @ (b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * eye)
        V = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * eye
    elif m == 9:
        U = A @ (b[9] * A8 + b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * eye)
        V = b[8] * A8 + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * eye
    else:
        sc = 2.0 ** -s
        As, A2, A4, A6 = A * sc, A2 * sc ** 2, A4 * sc ** 4, A6 * sc ** 6
        U2 = A6 @ (b[13] * A6 + b[11] * A4 + b[9] * A2)
        U = As @ (U2 + b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * eye)
        V2 = A6 @ (b[12] * A6 + b[10] * A4 + b[8] * A2)
        V = V2 + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * eye
    X = solve(V - U, V + U)
    for _ in range(s):
        X = X @ X
    return (X, m, s) if info else X
