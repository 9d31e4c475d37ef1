"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the script-level helpers
`detrend`, `forecast` (+ nested `MLII`) that every reference script re-defines.

Follows north/June1st.py:179-194 (detrend), :208-279 (forecast) and :235-257 (MLII); the retrospective
variants (north/retrospective_forecasts/June1st_retro.py:178-195, :210-291) differ only in year
bookkeeping and `.round(3)`.  The arithmetic is delegated to the same third-party routines the
reference calls (`scipy.stats.linregress`, `scipy.stats.pearsonr`, `np.cov`, `scipy.linalg.expm`,
`np.linalg.cholesky/solve/multi_dot`), so results are the reference's under this image's
numpy 2.3.5 / scipy 1.18.1.  Pinned by `tests/golden/make_golden.py`, which AST-lifts the reference's own
`detrend`/`forecast` function bodies and records their outputs (un-rounded) as fixtures.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm
from scipy.stats import linregress, pearsonr


def detrend(data):
    """north/June1st.py:179-194.  (X,Y,T) -> residuals (X,Y,T), trend (X,Y,2)=[slope,intercept]."""
    X, Y, T = data.shape
    detrended = np.zeros(data.shape) * np.nan
    trend = np.zeros((X, Y, 2)) * np.nan
    t = np.arange(T)
    for i in range(X):
        for j in range(Y):
            if ~np.isnan(data[i, j, :]).all():
                reg = linregress(t, data[i, j, :])
                lineT = (reg[0] * t) + reg[1]
                trend[i, j, 0] = reg[0]
                trend[i, j, 1] = reg[1]
                detrended[i, j, :] = data[i, j, :] - lineT
    return detrended, trend


def select_predictors(y, anoms_sic, anoms_sst=None, rule="pos", alpha=0.05):
    """Predictor loop, north/June1st.py:216-224 / north/August1st.py:174-181.
    rule: 'pos' (r>0), 'all', 'pos_sig' ((r>0) & (p/2<alpha)).  SST series enter negated when r<0.
    Returns the (n+1, Np) matrix (rows = time, last row = test year)."""
    cols = []
    for area in anoms_sic:
        r, p = pearsonr(y, anoms_sic[area][:-1])
        if rule == "all":
            cols.append(anoms_sic[area])
        elif rule == "pos":
            if r > 0:
                cols.append(anoms_sic[area])
        elif rule == "pos_sig":
            if (r > 0) & (p / 2 < alpha):
                cols.append(anoms_sic[area])
        else:
            raise ValueError(rule)
    if anoms_sst is not None:
        for area in anoms_sst:
            r, p = pearsonr(y, anoms_sst[area][:-1])
            if r < 0:
                cols.append(-anoms_sst[area])
    return np.asarray(cols).T


def design(Xfull, zscore):
    """north/June1st.py:226-233: optional column z-score over all n+1 rows, split, Laplacian prior."""
    X = Xfull
    if zscore:
        X = (X - np.mean(X, 0)) / np.std(X, 0)
    Xs = np.asarray([X[-1, :]])
    X = X[:-1, :]
    M = np.abs(np.cov(X, rowvar=False, bias=True))
    np.fill_diagonal(M, 0)            # one predictor: np.cov is 0-d and this raises ValueError, as in the reference
    np.fill_diagonal(M, -np.sum(M, axis=0))
    return X, Xs, M


def gp_fit_predict(X, Xs, y, M, ell, sig):
    """north/June1st.py:263-277.  y is (n,1).  Returns dict with fmean, fvar, sigma_f, nlml."""
    n = len(y)
    S_t = expm(ell * M)
    K_t = np.linalg.multi_dot([X, S_t, X.T]) + np.eye(n) * sig
    L_t = np.linalg.cholesky(K_t)
    A_t = np.linalg.solve(L_t.T, np.linalg.solve(L_t, y))
    sf = (np.dot(y.T, A_t) / n)[0][0]
    sn = sf * sig
    S = sf * expm(ell * M)
    L = np.linalg.cholesky(np.linalg.multi_dot([X, S, X.T]) + np.eye(n) * sn)
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
    KXXs = np.linalg.multi_dot([X, S, Xs.T])
    KXsXs = np.linalg.multi_dot([Xs, S, Xs.T]) + sn
    v = np.linalg.solve(L, KXXs)
    fmean = np.dot(KXXs.T, alpha)[0][0]
    fvar = (KXsXs - np.dot(v.T, v))[0][0]
    nlml = (np.dot(y.T, alpha) / 2 + np.log(L.diagonal()).sum() + n * np.log(2 * np.pi) / 2)[0][0]
    # cond: 2-norm condition number of the kernel matrix both Cholesky solves see (K = sf * K_t has the same one);
    # not a reference output -- the parity tests scale their tolerance with it
    return dict(fmean=fmean, fvar=fvar, sigma_f=sf, nlml=nlml, cond=float(np.linalg.cond(K_t)))


def mlii(theta, X, y, M):
    """The nested MLII, north/June1st.py:235-257 (negative log marginal likelihood + gradient as written)."""
    n = len(y)
    ell = np.exp(theta[0])
    sig = np.exp(theta[1])
    try:
        S_t = expm(ell * M)
        L_t = np.linalg.cholesky(np.linalg.multi_dot([X, S_t, X.T]) + np.eye(n) * sig)
        A_t = np.linalg.solve(L_t.T, np.linalg.solve(L_t, y))
        sf = (np.dot(y.T, A_t) / n)[0][0]
        sn = sf * sig
        S = sf * expm(ell * M)
        L = np.linalg.cholesky(np.linalg.multi_dot([X, S, X.T]) + np.eye(n) * sn)
        a = np.linalg.solve(L.T, np.linalg.solve(L, y))
        nlML = np.dot(y.T, a) / 2 + np.log(L.diagonal()).sum() + n * np.log(2 * np.pi) / 2
        dKdl = np.linalg.multi_dot([X, np.dot(M, S), X.T]) + np.eye(n) * sn
        dKds = np.linalg.multi_dot([X, S, X.T]) + np.eye(n) * sf
        g1 = ((np.trace(np.linalg.solve(L.T, np.linalg.solve(L, dKdl))) / 2
               - np.linalg.multi_dot([a.T, dKdl, a]) / 2))[0][0]
        g2 = ((np.trace(np.linalg.solve(L.T, np.linalg.solve(L, dKds))) / 2
               - np.linalg.multi_dot([a.T, dKds, a]) / 2))[0][0]
    except (np.linalg.LinAlgError, ValueError, OverflowError):
        nlML = np.inf
        g1 = np.inf
        g2 = np.inf
    return np.squeeze(nlML), np.asarray([g1, g2])


def forecast_one(y, anoms_sic, anoms_sst, rule, alpha, zscore, ell, sig, slope, icpt, t_index):
    """One (year, region) forecast: selection -> design -> fit/predict -> re-trend
    (north/retrospective_forecasts/June1st_retro.py:221-286 without the `.round(3)`)."""
    y = np.asarray(y, dtype=np.float64)
    Xfull = select_predictors(y, anoms_sic, anoms_sst, rule, alpha)
    X, Xs, M = design(Xfull, zscore)
    out = gp_fit_predict(X, Xs, y[:, None], M, ell, sig)
    out["fmean_rt"] = out["fmean"] + (slope * t_index + icpt)
    out["n_pred"] = Xfull.shape[1]
    return out
