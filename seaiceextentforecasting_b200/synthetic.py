"""Synthetic NSIDC-shaped inputs (no network access: the real SIC/SST/SIE files cannot be downloaded).

Shapes follow the reference scripts:
  * north SIC grid 57x57 (`make_npstere_grid(65,360,1e5)`, north/June1st.py:80), `latlon_grid=False`, weight `psar`
  * SST grid 26x90 lat-lon (`np.arange(90,38,-2)` x `np.arange(-180,180,4)`, north/June1st.py:166), `latlon_grid=True`
  * south SIC grid 81x81 (`make_npstere_grid(-55,180,1e5)`, south/February1st.py:79)
  * 25 km grid 448x304 (north/June1st.py:74-75)
Fields are a sum of Gaussian-blob spatial modes with random time amplitudes plus white noise and a
linear trend, clipped to [0,1] like a concentration; land is NaN (outside a disc plus random cells), so
every field has at least one all-NaN cell (ComplexNetworks.py:50-51 needs one).  Saturated cells are
exactly 0.0 or 1.0, so constant series detrend to exactly 0 and drop out of the node mask like they do in
the reference (ComplexNetworks.py:32).

Everything is numpy + a seeded `default_rng`; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

FIRST_YEAR = 1979  # north/retrospective_forecasts/June1st_retro.py:180 (`n = year-1979+1`)


def make_field(X, Y, T, seed, n_modes=None, noise=0.6, latlon=False, land_frac=0.05,
               saturate=True, blob=(2.0, 6.0)):
    """(X,Y,T) float64 field with NaN land; returns (data, amps) with amps (n_modes,T)."""
    rng = np.random.default_rng(seed)
    if n_modes is None:
        n_modes = max(6, int(round(40 * X * Y / (57.0 * 57.0))))
    xx, yy = np.meshgrid(np.arange(X, dtype=np.float64), np.arange(Y, dtype=np.float64), indexing="ij")
    cx = rng.uniform(0, X, n_modes)
    cy = rng.uniform(0, Y, n_modes)
    s = rng.uniform(blob[0], blob[1], n_modes) * max(1.0, min(X, Y) / 57.0) ** 0.5
    amps = rng.standard_normal((n_modes, T))
    field = np.zeros((X, Y, T))
    for m in range(n_modes):
        dx = xx - cx[m]
        dy = yy - cy[m]
        if latlon:  # periodic in longitude (second axis)
            dy = (dy + Y / 2.0) % Y - Y / 2.0
        blob = np.exp(-(dx * dx + dy * dy) / (2.0 * s[m] * s[m]))
        field += blob[:, :, None] * amps[m][None, None, :]
    field += noise * rng.standard_normal((X, Y, T))
    trend = -0.004 * np.arange(T)
    if saturate:
        # concentration-like: mid-pack cells vary, a rim of cells saturates at 0 or 1
        base = 0.5 + 0.35 * np.cos(np.pi * np.hypot(xx - X / 2.0, yy - Y / 2.0) / (0.6 * max(X, Y)))
        field = np.clip(base[:, :, None] + 0.22 * field + trend[None, None, :], 0.0, 1.0)
    else:
        field = field + trend[None, None, :]
    if latlon:
        land = rng.uniform(size=(X, Y)) < (land_frac * 4)
        land[0, :] = True  # a NaN row (pole / below min_lat) guarantees the sentinel cell
    else:
        r = np.hypot(xx - (X - 1) / 2.0, yy - (Y - 1) / 2.0)
        land = (r > 0.48 * min(X, Y)) | (rng.uniform(size=(X, Y)) < land_frac)
        land[0, 0] = True
    field[land] = np.nan
    return field, amps


def make_psar(X, Y, seed=0):
    """Polar-stereographic cell-area weights, `psar = 16*griddata(psa)` in the reference
    (north/September1st.py:137): smooth, positive, ~1e4 km^2."""
    xx, yy = np.meshgrid(np.arange(X, dtype=np.float64), np.arange(Y, dtype=np.float64), indexing="ij")
    r2 = (xx - (X - 1) / 2.0) ** 2 + (yy - (Y - 1) / 2.0) ** 2
    return 16.0 * (664.4 - 281.7 * r2 / r2.max())


def make_lat_grid(X, Y):
    """SST latitude grid, `np.meshgrid(np.arange(-180,180,4), np.arange(90,38,-2))[1]` (north/June1st.py:166)
    generalised to X rows."""
    return np.tile((90.0 - 2.0 * np.arange(X, dtype=np.float64))[:, None], (1, Y))


def make_sie(field, T, seed, n_regions=3, noise=0.35, lag=0):
    """Three regional extent series (million km^2, 3 d.p. like the Sea Ice Index files read at
    north/retrospective_forecasts/June1st_retro.py:52-54).  Region 0 follows the domain-mean concentration,
    regions 1 and 2 follow two sectors, so that several network areas correlate with each target (the
    reference's `forecast()` crashes when fewer than two predictors pass its selection rule).  `lag=1`: the
    target of year t follows the field of year t-1 (south January/December scripts forecast from the previous
    year's network, south/retrospective_forecasts/January1st_retro.py:175-178)."""
    rng = np.random.default_rng(seed + 7919)
    X, Y = field.shape[:2]
    xx, yy = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    sectors = [np.ones((X, Y), dtype=bool), (xx < 0.55 * X) & (yy < 0.6 * Y), (xx > 0.4 * X) & (yy > 0.35 * Y)]
    out = []
    for k in range(n_regions):
        with np.errstate(invalid="ignore"):
            m = np.nanmean(field[sectors[k % 3]][:, :T], axis=0)
        if lag:
            m = np.concatenate([m[:lag], m[:-lag]])
        m = m - m.mean()
        m = m / (m.std() + 1e-12)
        a = [6.5, 0.6, 0.5][k % 3]
        b = [0.08, 0.012, 0.010][k % 3]
        sc = [0.35, 0.08, 0.07][k % 3]
        y = a - b * np.arange(T) + sc * m + noise * sc * rng.standard_normal(T)
        out.append(np.round(y, 3))
    return out
