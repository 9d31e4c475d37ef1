"""Pins the oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by make_golden.py from
/root/reference).  Networks: bit-exact V (keys, membership, list order), nodes, tau, anomaly, strengthmap.
Sweeps: the AST-lifted reference detrend/networks/forecast/MLII of every retrospective script vs oracle/sweep.py."""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import gp as ogp
from oracle import sweep as osweep
from oracle.network import Network
from seaiceextentforecasting_b200.config import CONFIGS, RULE_ALL, RULE_POS, RULE_POS_SIG

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = sorted(glob.glob(os.path.join(GOLD, "network_*.npz")))
SWEEPS = sorted(glob.glob(os.path.join(GOLD, "sweep_*.npz")))


def unpack_V(g, suffix=""):
    keys, lens, cells = g["V_keys" + suffix], g["V_lens" + suffix], g["V_cells" + suffix]
    V, off = {}, 0
    for k, ln in zip(keys, lens):
        V[int(k)] = [[int(a), int(b)] for a, b in cells[off:off + ln]]
        off += ln
    return V


def test_fixtures_present():
    assert len(NETS) == 6 and len(SWEEPS) == 7


@pytest.mark.parametrize("path", NETS, ids=[os.path.basename(p) for p in NETS])
def test_network_oracle_equals_reference(path):
    g = np.load(path)
    dt, trend = ogp.detrend(g["raw"])
    assert np.array_equal(dt, g["dt"], equal_nan=True)            # detrend(): bit-exact vs north/June1st.py:179-194
    assert np.array_equal(trend, g["trend"], equal_nan=True)
    latlon = bool(g["latlon"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = Network(data=g["dt"])
        Network.tau(o, 0.01)
        Network.area_level(o, latlon_grid=latlon)
        if latlon:
            Network.intra_links(o, lat=g["weight"])
        else:
            Network.intra_links(o, area=g["weight"])
    assert np.array_equal(o.nodes, g["nodes"]) and o.nodes.dtype == g["nodes"].dtype
    assert o.tau == float(g["tau"])
    V = unpack_V(g)
    assert list(o.V.keys()) == list(V.keys())
    assert all(o.V[k] == V[k] for k in V)
    assert o.V is o.A and bool(g["V_is_A"])
    assert np.array_equal(np.array([o.anomaly[k] for k in o.V]), g["anomaly"])
    assert np.array_equal(np.array([o.links[k] for k in o.V], dtype=float), g["links"])
    assert np.array_equal(np.array([o.strength[k] for k in o.V]), g["strength"])
    assert np.array_equal(o.strengthmap, g["strengthmap"], equal_nan=True)
    assert np.array_equal(o.corrs[:4], g["corrs_rows"], equal_nan=True)


@pytest.mark.parametrize("path", SWEEPS, ids=[os.path.basename(p) for p in SWEEPS])
def test_sweep_oracle_equals_reference(path):
    g = np.load(path)
    name = os.path.basename(path)[len("sweep_"):-len(".npz")]
    cfg = CONFIGS[name]
    fmin, fmax = int(g["fmin"]), int(g["fmax"])
    sie = {r: g["sie"][i] for i, r in enumerate(cfg.regions)}
    for r in cfg.regions:                                          # read_SIE tables (June1st_retro.py:58-69)
        dt, tr = osweep.sie_tables(sie[r], fmin, fmax)
        assert np.array_equal(dt, g["siedt_" + r]) and np.array_equal(tr, g["sietrend_" + r])
    out = osweep.retro_sweep([cfg], {name: g["sic"]}, sie, fmin, fmax, g["psar"],
                             g["sst"] if cfg.use_sst else None, g["sst_lat"] if cfg.use_sst else None)[name]
    for year, V in out["V"].items():
        assert V == unpack_V(g, f"_{year}"), year                 # every network of the sweep, bit-exact
    for r in cfg.regions:
        for suf in ("_fmean", "_fvar", "_fmean_rt"):
            ref = g["raw_" + r + suf]
            np.testing.assert_allclose(out[r + suf], ref, rtol=1e-12, atol=1e-13)
            assert np.array_equal(np.round(out[r + suf], 3) if suf != "_fmean_rt" else g["rnd_" + r + suf],
                                  g["rnd_" + r + suf])


@pytest.mark.parametrize("path", SWEEPS, ids=[os.path.basename(p) for p in SWEEPS])
def test_mlii_oracle_equals_reference(path):
    g = np.load(path)
    name = os.path.basename(path)[len("sweep_"):-len(".npz")]
    cfg = CONFIGS[name]
    fmin, fmax = int(g["fmin"]), int(g["fmax"])
    year = fmax
    ny = year - 1 if cfg.prev_year_network else year
    _, anoms, _ = osweep.build_network(g["sic"], ny, False, g["psar"])
    sst_anoms = None
    if cfg.use_sst:
        _, sst_anoms, _ = osweep.build_network(g["sst"], year, True, g["sst_lat"])
    row = year - (fmin - 1) - 1
    sdt = g["siedt_" + cfg.regions[0]]
    y = sdt[row, 1:year - 1979] if cfg.prev_year_network else sdt[row, 0:year - 1979]
    rule = {RULE_POS: "pos", RULE_ALL: "all", RULE_POS_SIG: "pos_sig"}[cfg.rule[0]]
    X, Xs, M = ogp.design(ogp.select_predictors(y, anoms, sst_anoms, rule, cfg.alpha), cfg.zscore)
    nl, grad = ogp.mlii(g["mlii_theta"], X, y[:, None], M)
    assert float(nl) == float(g["mlii_nl"])
    assert np.array_equal(grad, g["mlii_grad"])
