"""ORACLE (test infrastructure, NOT product code) -- the retrospective sweep of the reference scripts
(north/retrospective_forecasts/*_retro.py, south/retrospective_forecasts/*_retro.py: read_SIE tables ->
detrend -> networks -> forecast) restated over in-memory synthetic inputs, one (config, year) job at a time so a
bounded sample of the full workload can be timed as the CPU baseline and fanned over host cores.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs import this.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.stats import linregress

from . import gp as ogp
from .network import Network

FIRST_YEAR = 1979
_RULES = {0: "pos", 1: "all", 2: "pos_sig"}


def sie_tables(sie, fmin, fmax):
    """read_SIE de-trending loop, north/retrospective_forecasts/June1st_retro.py:58-69."""
    trend = np.zeros((fmax - (fmin - 1) + 1, 2))
    dt = np.zeros((fmax - (fmin - 1) + 1, fmax - FIRST_YEAR + 1))
    sie = np.asarray(sie, dtype=np.float64)
    for year in range(fmin - 1, fmax + 1):
        n = year - FIRST_YEAR + 1
        reg = linregress(np.arange(n), sie[range(n)])
        lineT = (reg[0] * np.arange(n)) + reg[1]
        trend[year - (fmin - 1), 0] = reg[0]
        trend[year - (fmin - 1), 1] = reg[1]
        dt[year - (fmin - 1), range(n)] = sie[range(n)] - lineT
    return dt.round(3), trend


def build_network(field, year, latlon, weight, significance=0.01):
    """detrend + networks for one year window (June1st_retro.py:178-208) -> (V, anomaly)."""
    n = year - FIRST_YEAR + 1
    dt, _ = ogp.detrend(field[:, :, :n])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = Network(data=dt)
        Network.tau(net, significance, keep_corrs=False)
        Network.area_level(net, latlon_grid=latlon)
        if latlon:
            Network.intra_links(net, lat=weight)
        else:
            Network.intra_links(net, area=weight)
    return net.V, net.anomaly, net.tau


FAILS = (IndexError, ValueError, np.linalg.LinAlgError)   # what forecast() raises: no / one predictor, non-SPD K


def run_job(cfg, year, sic_field, psar, sie_dt, sie_trend, fmin, sst_field=None, sst_lat=None, record_failures=False):
    """All three regional forecasts of one (config, target year): the network build(s) plus forecast().
    `cfg` is a seaiceextentforecasting_b200.config.ForecastConfig (plain data).  `record_failures`: a region on
    which forecast() raises (the reference aborts the whole script there) becomes a record of NaNs with
    `failed = <exception name>` instead of propagating."""
    ny = year - 1 if cfg.prev_year_network else year
    V, anoms, tau = build_network(sic_field, ny, False, psar)
    sst_anoms = None
    if cfg.use_sst:
        _, sst_anoms, _ = build_network(sst_field, year, True, sst_lat)
    out = []
    row = year - (fmin - 1) - 1
    for k, reg in enumerate(cfg.regions):
        if cfg.prev_year_network:
            y = sie_dt[reg][row, 1:year - FIRST_YEAR]
        else:
            y = sie_dt[reg][row, 0:year - FIRST_YEAR]
        slope, icpt = sie_trend[reg][row]
        try:
            r = ogp.forecast_one(y, anoms, sst_anoms, _RULES[cfg.rule[k]], cfg.alpha, cfg.zscore, cfg.ell[k], cfg.sig[k],
                                 slope, icpt, year - FIRST_YEAR)
            r["failed"] = None
        except FAILS as e:
            if not record_failures:
                raise
            r = dict(fmean=np.nan, fvar=np.nan, sigma_f=np.nan, nlml=np.nan, cond=np.nan, n_pred=-1,
                     failed=type(e).__name__)
        lineT = (np.arange(year - FIRST_YEAR + 1) * slope) + icpt
        r["fmean_rt"] = r["fmean"] + lineT[-1]
        r["n_areas"] = len(V)
        out.append(r)
    return out, V


def retro_sweep(cfgs, sic_fields, sie, fmin, fmax, psar, sst_field=None, sst_lat=None, record_failures=False):
    """-> {config: {region+'_fmean'|'_fvar'|'_fmean_rt': array(years)}} un-rounded, plus V per network year (and, with
    `record_failures`, region+'_failed': list of exception names / None and region+'_cond')."""
    out = {}
    for cfg in cfgs:
        tables = {reg: sie_tables(sie[reg], fmin, fmax) for reg in cfg.regions}
        sie_dt = {reg: tables[reg][0] for reg in cfg.regions}
        sie_trend = {reg: tables[reg][1] for reg in cfg.regions}
        g = out.setdefault(cfg.name, {"V": {}})
        for reg in cfg.regions:
            for suf in ("_fmean", "_fvar", "_fmean_rt", "_cond"):
                g[reg + suf] = np.zeros(fmax - fmin + 1)
            g[reg + "_failed"] = [None] * (fmax - fmin + 1)
        for year in range(fmin, fmax + 1):
            res, V = run_job(cfg, year, sic_fields[cfg.name], psar, sie_dt, sie_trend, fmin, sst_field, sst_lat,
                             record_failures)
            g["V"][year - 1 if cfg.prev_year_network else year] = V
            for k, reg in enumerate(cfg.regions):
                g[reg + "_fmean"][year - fmin] = res[k]["fmean"]
                g[reg + "_fvar"][year - fmin] = res[k]["fvar"]
                g[reg + "_fmean_rt"][year - fmin] = res[k]["fmean_rt"]
                g[reg + "_cond"][year - fmin] = res[k]["cond"]
                g[reg + "_failed"][year - fmin] = res[k]["failed"]
    return out
