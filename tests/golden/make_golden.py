"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes next to this file

The reference has no tests or golden vectors of its own, so the oracle is pinned against
outputs of the reference itself:
  * network_<case>.npz  : /root/reference/ComplexNetworks.py (class Network) imported as-is and run on seeded
                          synthetic de-trended grids -> nodes, tau, V (keys / cell lists / order), anomaly, links,
                          strength, strengthmap, a slice of `corrs`
  * detrend.npz         : `detrend()` AST-lifted from north/June1st.py:179-194
  * sweep_<script>.npz  : `detrend`/`networks`/`forecast`/`skill` AST-lifted from each retrospective script and run
                          end to end on a small synthetic data set with the `.round(3)` calls on the GP outputs
                          stripped (`fmean/fvar/fmean_rt` un-rounded), plus the nested `MLII` evaluated at its x0
Nothing is copied from the reference: functions are parsed out of the files where they lie and exec'd.
"""
import ast
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.simplefilter("ignore")

import ComplexNetworks as REFCN  # noqa: E402  (the reference module)
from scipy.linalg import expm  # noqa: E402
from scipy.stats import linregress, pearsonr  # noqa: E402

from seaiceextentforecasting_b200 import synthetic as syn  # noqa: E402

REFCN.CN = REFCN   # `from ComplexNetworks import CN` (north/June1st.py:197)
_alias = types.ModuleType("CNs_backup.backups")
_alias.CN_forecast = REFCN
sys.modules["CNs_backup"] = types.ModuleType("CNs_backup")
sys.modules["CNs_backup.backups"] = _alias   # north/retrospective_forecasts/June1st_retro.py:198


class _StripRound3(ast.NodeTransformer):
    """x.round(3) -> x   (only inside forecast(): GP outputs un-rounded)"""

    def visit_Call(self, node):
        self.generic_visit(node)
        if (isinstance(node.func, ast.Attribute) and node.func.attr == "round" and len(node.args) == 1
                and isinstance(node.args[0], ast.Constant) and node.args[0].value == 3):
            return node.func.value
        return node


def lift(path, names, strip_round_in=("forecast",)):
    """Parse the function definitions `names` out of a reference script and return them exec'd in a namespace."""
    tree = ast.parse(open(path, encoding="utf-8").read())
    ns = {"np": np, "linregress": linregress, "pearsonr": pearsonr, "expm": expm, "CN": REFCN}
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            if node.name in strip_round_in:
                node = _StripRound3().visit(node)
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), ns)
    return ns


def lift_nested_mlii(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "forecast":
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef) and sub.name == "MLII":
                    mod = ast.Module(body=[sub], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = {"np": np, "expm": expm}
                    exec(compile(mod, path, "exec"), ns)
                    return ns
    raise RuntimeError("MLII not found in " + path)


def pack_V(V):
    keys = np.array(list(V.keys()), dtype=np.int64)
    lens = np.array([len(V[k]) for k in V], dtype=np.int64)
    cells = np.array([c for k in V for c in V[k]], dtype=np.int64).reshape(-1, 2)
    return keys, lens, cells


# ------------------------------------------------------------------------------------------- networks
NETWORK_CASES = {  # name: X, Y, T, latlon, seed
    "a14": (14, 14, 20, False, 1), "b16": (16, 12, 9, False, 2), "c10ll": (10, 24, 30, True, 3),
    "d12ll": (12, 30, 12, True, 4), "e20": (20, 20, 42, False, 5), "f24": (24, 24, 7, False, 6),
}


def make_networks():
    june = lift(os.path.join(REF, "north/June1st.py"), ("detrend",))
    for name, (X, Y, T, latlon, seed) in NETWORK_CASES.items():
        data, _ = syn.make_field(X, Y, T, seed, latlon=latlon)
        ds = {"data": data}
        june["detrend"](ds)
        dt = ds["dt"]
        net = REFCN.Network(data=dt)
        REFCN.Network.tau(net, 0.01)
        REFCN.Network.area_level(net, latlon_grid=latlon)
        if latlon:
            lat = syn.make_lat_grid(X, Y)
            REFCN.Network.intra_links(net, lat=lat)
        else:
            area = syn.make_psar(X, Y)
            REFCN.Network.intra_links(net, area=area)
        keys, lens, cells = pack_V(net.V)
        np.savez_compressed(
            os.path.join(HERE, f"network_{name}.npz"), raw=data, dt=dt, trend=ds["trend"], latlon=latlon,
            weight=(lat if latlon else area), nodes=net.nodes, tau=net.tau, V_keys=keys, V_lens=lens, V_cells=cells,
            anomaly=np.array([net.anomaly[k] for k in net.V]), links=np.array([net.links[k] for k in net.V], dtype=float),
            strength=np.array([net.strength[k] for k in net.V]), strengthmap=net.strengthmap,
            corrs_rows=net.corrs[:4].copy(), V_is_A=(net.V is net.A))
        print("network", name, "N", net.nodes.shape[1], "tau", net.tau, "areas", len(net.V))


# ------------------------------------------------------------------------------------------- sweeps
SWEEPS = {
    "north_june": ("north/retrospective_forecasts/June1st_retro.py", True, False),
    "north_july": ("north/retrospective_forecasts/July1st_retro.py", False, False),
    "north_august": ("north/retrospective_forecasts/August1st_retro.py", False, False),
    "north_september": ("north/retrospective_forecasts/September1st_retro.py", False, False),
    "south_february": ("south/retrospective_forecasts/February1st_retro.py", False, False),
    "south_january": ("south/retrospective_forecasts/January1st_retro.py", False, True),
    "south_december": ("south/retrospective_forecasts/December1st_retro.py", False, True),
}
FMIN, FMAX = 1991, 1994        # windows T = 13..16 (n = 12..15)
GX, GY, SX, SY = 15, 15, 8, 18


def sweep_inputs(seed, lag=0):
    Tfull = FMAX - 1979 + 1
    sic, amps = syn.make_field(GX, GY, Tfull, seed, n_modes=50, noise=0.5, blob=(1.0, 2.5))
    sst, _ = syn.make_field(SX, SY, Tfull, seed + 1, latlon=True, saturate=False, n_modes=20, noise=0.5,
                            blob=(1.0, 2.5))
    sie = syn.make_sie(sic, Tfull, seed, lag=lag)
    return sic, sst, sie


def ref_read_sie_tables(sie_list, regions, fmin, fmax):
    """The loop of read_SIE (June1st_retro.py:58-69) applied to in-memory series (the download part is I/O)."""
    SIEs = {r: s for r, s in zip(regions, sie_list)}
    SIEs_dt, SIEs_trend = {}, {}
    for tag in SIEs:
        trend = np.zeros((fmax - (fmin - 1) + 1, 2))
        dt = np.zeros((fmax - (fmin - 1) + 1, fmax - 1979 + 1))
        for year in range(fmin - 1, fmax + 1):
            n = year - 1979 + 1
            reg = linregress(np.arange(n), SIEs[tag][range(n)])
            lineT = (reg[0] * np.arange(n)) + reg[1]
            trend[year - (fmin - 1), 0] = reg[0]
            trend[year - (fmin - 1), 1] = reg[1]
            dt[year - (fmin - 1), range(n)] = SIEs[tag][range(n)] - lineT
        SIEs_trend[tag] = trend
        SIEs_dt[tag] = dt.round(3)
    return SIEs, SIEs_dt, SIEs_trend


def make_sweeps():
    for si, (name, (rel, use_sst, prev_year)) in enumerate(SWEEPS.items()):
        path = os.path.join(REF, rel)
        ns = lift(path, ("detrend", "networks", "forecast", "skill"))
        sic, sst, sie = sweep_inputs(100 + si, lag=1 if prev_year else 0)
        regions = ["Pan-Arctic", "Beaufort", "Chukchi"] if name.startswith("north") else ["Pan-Antarctic", "Ross", "Weddell"]
        SIEs, SIEs_dt, SIEs_trend = ref_read_sie_tables(sie, regions, FMIN, FMAX)
        SIC = {"data": sic, "psar": syn.make_psar(GX, GY)}
        ns.update(SIEs=SIEs, SIEs_dt=SIEs_dt, SIEs_trend=SIEs_trend, SIC=SIC)
        n0, n1 = (FMIN - 1, FMAX - 1) if prev_year else (FMIN, FMAX)     # south January1st_retro.py:284-287
        ns["detrend"](SIC, n0, n1)
        if name == "north_june":
            ns["networks"](SIC, n0, n1, latlon=False)
            SST = {"data": sst, "lat": syn.make_lat_grid(SX, SY)}
            ns["SST"] = SST
            ns["detrend"](SST, FMIN, FMAX)
            ns["networks"](SST, FMIN, FMAX, latlon=True)
        else:
            ns["networks"](SIC, n0, n1)
        GPR = ns["forecast"](FMIN, FMAX)
        ns["GPR"] = GPR
        # the rounded variant for skill(): re-run forecast un-stripped
        ns_r = lift(path, ("forecast", "skill"), strip_round_in=())
        ns_r.update({k: ns[k] for k in ("SIEs", "SIEs_dt", "SIEs_trend", "SIC")})
        if use_sst:
            ns_r["SST"] = ns["SST"]
        GPR_r = ns_r["forecast"](FMIN, FMAX)
        ns_r["GPR"] = GPR_r
        skill_rt, skill_dt, dt_obs = ns_r["skill"](FMIN, FMAX)
        out = dict(sic=sic, sie=np.array(sie), fmin=FMIN, fmax=FMAX, psar=SIC["psar"],
                   skill_rt=np.array(skill_rt), skill_dt=np.array(skill_dt))
        if use_sst:
            out.update(sst=sst, sst_lat=ns["SST"]["lat"])
        for r in regions:
            for suf in ("_fmean", "_fvar", "_fmean_rt"):
                out["raw_" + r + suf] = GPR[r + suf]
                out["rnd_" + r + suf] = GPR_r[r + suf]
            out["siedt_" + r] = SIEs_dt[r]
            out["sietrend_" + r] = SIEs_trend[r]
        # area counts per network year (cheap structural check)
        yrs = range(n0, n1 + 1)
        out["n_areas"] = np.array([len(SIC["nodes_" + str(y)]) for y in yrs])
        for y in yrs:
            keys, lens, cells = pack_V(SIC["nodes_" + str(y)])
            out[f"V_keys_{y}"], out[f"V_lens_{y}"], out[f"V_cells_{y}"] = keys, lens, cells
        # nested MLII at x0 for region 0 of the last year, rebuilt from the same X, y, M the script builds
        year = FMAX
        ml = lift_nested_mlii(path)
        from oracle import gp as og   # only to rebuild X/M inputs for the reference's MLII; outputs are the reference's
        ny = year - 1 if prev_year else year
        if prev_year:
            y = SIEs_dt[regions[0]][year - (FMIN - 1) - 1, range(1, year - 1979)]
        else:
            y = SIEs_dt[regions[0]][year - (FMIN - 1) - 1, range(year - 1979)]
        from seaiceextentforecasting_b200.config import CONFIGS, RULE_ALL, RULE_POS, RULE_POS_SIG
        cfg = CONFIGS[name]
        rule = {RULE_POS: "pos", RULE_ALL: "all", RULE_POS_SIG: "pos_sig"}[cfg.rule[0]]
        Xfull = og.select_predictors(y, SIC["anoms_" + str(ny)], ns["SST"]["anoms_" + str(year)] if use_sst else None,
                                     rule, cfg.alpha)
        X, Xs, M = og.design(Xfull, cfg.zscore)
        ml.update(X=X, Xs=Xs, y=y[:, None], n=len(y), M=M)
        theta = [np.log(cfg.ell[0]), np.log(cfg.sig[0])]
        nl, grad = ml["MLII"](theta)
        out.update(mlii_theta=np.array(theta), mlii_nl=float(nl), mlii_grad=np.array(grad, dtype=float))
        np.savez_compressed(os.path.join(HERE, f"sweep_{name}.npz"), **out)
        print("sweep", name, "areas", out["n_areas"], "fmean0", GPR[regions[0] + "_fmean"], "mlii", nl)


if __name__ == "__main__":
    make_networks()
    make_sweeps()
