"""Host-side logic that needs no GPU: config table, year/row bookkeeping of the sweep, SIE tables, thresholds,
task sharding (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import pytest
from scipy import stats

from oracle import sweep as osweep
from seaiceextentforecasting_b200.config import CONFIGS, LS, NORTH_INITS, RULE_ALL, RULE_POS, RULE_POS_SIG, SS
from seaiceextentforecasting_b200.engine import pad_T, r_crit_pearson, r_crit_ttest
from seaiceextentforecasting_b200.forecast import GP_RESULT_DTYPE, SweepPlan, sie_detrend_tables


def test_config_table_matches_reference_scripts():
    c = CONFIGS
    assert c["north_june"].ell == (LS[16], LS[14], LS[12]) and c["north_june"].sig == (SS[1], SS[4], SS[6])
    assert c["north_june"].zscore and c["north_june"].use_sst and c["north_june"].rule == (RULE_POS,) * 3
    assert c["north_july"].ell[2] == 3.125433e+10 and c["north_july"].sig[2] == 40221.26298973
    assert c["north_august"].alpha == 0.08 and c["north_august"].rule == (RULE_ALL, RULE_POS_SIG, RULE_POS_SIG)
    assert c["north_september"].alpha == 0.05 and c["north_september"].ell == (LS[8], LS[9], LS[3])
    assert c["south_february"].sig == (SS[0], SS[11], SS[13])
    assert c["south_january"].prev_year_network and c["south_january"].alpha == 0.08
    assert c["south_december"].prev_year_network and c["south_december"].rule == (RULE_POS,) * 3


def test_thresholds_equivalent_to_reference_tests():
    rng = np.random.default_rng(0)
    for T in (7, 12, 30, 42):
        rc = r_crit_ttest(T, 0.01)
        R = rng.uniform(0, 1, 4000)
        df = T - 2
        P = stats.t.sf(R * np.sqrt(df / (1 - R ** 2)), df)          # ComplexNetworks.py:43-45
        assert np.array_equal(P < 0.01, R > rc)
    for n in (6, 10, 20, 41):
        for a in (0.05, 0.08):
            rc = r_crit_pearson(n, a)
            x = rng.standard_normal((300, n))
            y = rng.standard_normal(n)
            for row in x:
                r, p = stats.pearsonr(y, row)
                assert bool((r > 0) & (p / 2 < a)) == bool(r > 0 and r > rc)


def test_pad_T_is_conflict_free_stride():
    for T in range(3, 64):
        Tp = pad_T(T)
        assert Tp >= T and Tp % 8 == 4 and Tp - T < 8


def test_sie_tables_match_oracle_linregress():
    rng = np.random.default_rng(3)
    sie = np.round(6 - 0.08 * np.arange(42) + 0.3 * rng.standard_normal(42), 3)
    dt, tr = sie_detrend_tables(sie, 1985, 2020)
    odt, otr = osweep.sie_tables(sie, 1985, 2020)
    assert np.array_equal(dt, odt)                                   # rounded to 3 d.p. in both
    np.testing.assert_allclose(tr, otr, rtol=1e-12, atol=1e-14)


def _sie():
    rng = np.random.default_rng(0)
    return {r: np.round(rng.standard_normal(42), 3) for r in ("Pan-Arctic", "Beaufort", "Chukchi")}


def test_north_sweep_plan_counts():
    p = SweepPlan(NORTH_INITS, _sie(), 1985, 2020)
    assert len(p.jobs) == 144 and len(p.sst_years) == 36 and p.P == 432
    assert p.job_T.min() == 7 and p.job_T.max() == 42
    for (ci, k, year), pr in zip(p.prob_meta, p.prob):
        assert pr["n"] == year - 1979 and p.job_T[pr["job_sic"]] == pr["n"] + 1
        assert (pr["job_sst"] >= 0) == p.cfgs[ci].use_sst


def test_south_prev_year_plan():
    sie = {r: v for r, v in zip(("Pan-Antarctic", "Ross", "Weddell"), _sie().values())}
    p = SweepPlan(["south_january"], sie, 1985, 2020)
    for (ci, k, year), pr in zip(p.prob_meta, p.prob):
        assert pr["n"] == year - 1979 - 1                            # y drops 1979 (January1st_retro.py:175)
        assert p.jobs[pr["job_sic"]] == (0, year - 1)                # previous year's network
        assert p.job_T[pr["job_sic"]] == pr["n"] + 1


def test_shards_partition_the_sweep():
    full = SweepPlan(NORTH_INITS, _sie(), 1985, 2020)
    for world in (2, 4, 8):
        parts = [SweepPlan(NORTH_INITS, _sie(), 1985, 2020, rank=r, world=world) for r in range(world)]
        metas = [m for p in parts for m in p.prob_meta]
        assert sorted(metas) == sorted(full.prob_meta)
        loads = [int((p.job_T.astype(np.int64) ** 2).sum()) for p in parts]
        assert max(loads) <= 1.15 * min(loads)                       # balanced by window length


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from seaiceextentforecasting_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for members in (1, 3):                                           # 3 members: ranks own unequal problem counts
        plan = SweepPlan(NORTH_INITS, _sie(), 2015, 2020, rank=rank, world=world, members=members)
        raw = np.zeros(plan.P, dtype=GP_RESULT_DTYPE)
        for i, ((ci, k, year), m) in enumerate(zip(plan.prob_meta, plan.prob_member)):   # fake kernel output = problem id
            raw["fmean"][i] = m * 10000 + ci * 1000 + k * 100 + (year - 2000)
        out = parallel.gather_results(plan, raw)                     # tensor all-gather of the 80-byte records
        outs = out if members > 1 else [out]
        assert len(outs) == members
        for m, o in enumerate(outs):
            for ci, cfg in enumerate(plan.cfgs):
                for k, reg in enumerate(cfg.regions):
                    exp = np.array([m * 10000 + ci * 1000 + k * 100 + (y - 2000) for y in plan.years], dtype=float)
                    ok &= np.array_equal(o[cfg.name][reg + "_raw_fmean"], exp)
    # labels / node-series collectives (parallel.all_gather_networks, broadcast_networks) on CPU tensors
    import torch
    n_total, C, MA, T = 5, 12, 4, 6                                  # 5 networks over 2 ranks: rank 0 owns 0,2,4, rank 1 owns 1,3

    def fake_state(jobs):
        st = {}
        for k in parallel.NETWORK_TENSORS:
            shape = {"n_areas": (), "area_key": (MA,), "area_start": (MA + 1,), "area_cells": (C,), "label": (C,),
                     "status": (), "anomaly": (MA, T)}[k]
            dt = torch.float64 if k == "anomaly" else torch.int32
            st[k] = torch.stack([torch.full(shape, float(j * 10 + len(k)), dtype=dt) for j in jobs]) if jobs else \
                torch.zeros((0,) + shape, dtype=dt)
        return st

    mine = list(range(rank, n_total, world))
    full = parallel.all_gather_networks(fake_state(mine), n_total)
    for k in parallel.NETWORK_TENSORS:
        for j in range(n_total):
            ok &= bool((full[k][j] == j * 10 + len(k)).all())
    st = fake_state([0, 1]) if rank == 1 else {k: torch.zeros_like(v) for k, v in fake_state([0, 1]).items()}
    parallel.broadcast_networks(st, src=1)
    ref = fake_state([0, 1])
    ok &= all(bool(torch.equal(st[k], ref[k])) for k in parallel.NETWORK_TENSORS)
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_assembles_full_sweep():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_retro_csv_layout_matches_reference_script():
    """the CSV tables of June1st_retro.py:344-364, rebuilt from SweepPlan.assemble() output,
    against a literal restatement of those lines (prep/zip/DataFrame) on the same numbers."""
    import pandas as pd
    from seaiceextentforecasting_b200 import report
    from seaiceextentforecasting_b200.config import CONFIGS
    from seaiceextentforecasting_b200.forecast import GP_RESULT_DTYPE, SweepPlan
    rng = np.random.default_rng(5)
    T = 2020 - 1979 + 1
    sie = {r: np.round(6 - 0.05 * np.arange(T) + 0.4 * rng.standard_normal(T), 3) for r in CONFIGS["north_june"].regions}
    plan = SweepPlan(["north_june"], sie, 2010, 2020)
    raw = np.zeros(plan.P, dtype=GP_RESULT_DTYPE)
    raw["fmean"] = rng.standard_normal(plan.P)
    raw["fvar"] = rng.uniform(0.01, 0.5, plan.P)
    gpr = plan.assemble(raw)
    df_dt, df_rt = report.retro_tables(plan, gpr, "north_june")
    # --- literal restatement of June1st_retro.py:293-361 on the same arrays
    fmin, fmax = 2010, 2020
    GPR, SIEs, SIEs_dt = gpr["north_june"], plan.sie, plan.sie_dt
    regions = ['Pan-Arctic', 'Beaufort', 'Chukchi']
    skill_rt, skill_dt, dt_obs = [], [], []
    for k in range(3):
        dt = [SIEs_dt[regions[k]][t - (fmin - 1), t - 1979] for t in range(fmin, fmax + 1)]
        dt_obs.append(dt)
        forecast_rt = GPR[regions[k] + '_fmean_rt']
        obs_rt = SIEs[regions[k]][fmin - 1979:]
        a = np.mean((obs_rt - forecast_rt) ** 2)
        b = np.mean((obs_rt - np.nanmean(obs_rt)) ** 2)
        skill_rt.append((1 - (a / b)).round(3))
        c = np.mean((dt - GPR[regions[k] + '_fmean']) ** 2)
        d = np.mean((dt - np.nanmean(dt)) ** 2)
        skill_dt.append((1 - (c / d)).round(3))
    years = np.arange(fmin, fmax + 1).tolist()
    years.append('Skill')

    def prep(data, skill=None):
        if type(data) != list:
            data = data.tolist()
        data.append(skill if skill is not None else '')
        return data
    columns1 = ['Pan-Arctic$_o$', 'Pan-Arctic$_f$', 'Pan-Arctic$_f$ unc', 'Beaufort$_o$', 'Beaufort$_f$', 'Beaufort$_f$ unc',
                'Chukchi$_o$', 'Chukchi$_f$', 'Chukchi$_f$ unc']
    columns2 = ['Pan-Arctic$_o$', 'Pan-Arctic$_f$', 'Beaufort$_o$', 'Beaufort$_f$', 'Chukchi$_o$', 'Chukchi$_f$']
    data_dt = list(zip(prep(dt_obs[0]), prep(GPR['Pan-Arctic_fmean'], skill_dt[0]), prep(np.sqrt(GPR['Pan-Arctic_fvar']).round(3)),
                       prep(dt_obs[1]), prep(GPR['Beaufort_fmean'], skill_dt[1]), prep(np.sqrt(GPR['Beaufort_fvar']).round(3)),
                       prep(dt_obs[2]), prep(GPR['Chukchi_fmean'], skill_dt[2]), prep(np.sqrt(GPR['Chukchi_fvar']).round(3))))
    ref_dt = pd.DataFrame(data_dt, index=years, columns=columns1)
    data_rt = list(zip(prep(SIEs['Pan-Arctic'][fmin - 1979:]), prep(GPR['Pan-Arctic_fmean_rt'], skill_rt[0]),
                       prep(SIEs['Beaufort'][fmin - 1979:]), prep(GPR['Beaufort_fmean_rt'], skill_rt[1]),
                       prep(SIEs['Chukchi'][fmin - 1979:]), prep(GPR['Chukchi_fmean_rt'], skill_rt[2])))
    ref_rt = pd.DataFrame(data_rt, index=years, columns=columns2)
    assert df_dt.equals(ref_dt) and df_rt.equals(ref_rt)
    assert list(df_dt.index) == years and list(df_dt.columns) == columns1


def _tau_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from seaiceextentforecasting_b200.parallel import shard_tile_rows, tau_from_shards
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                 # same matrix on every rank
    N, rc = 700, 0.35
    Z = rng.standard_normal((N, 30))
    Z -= Z.mean(1, keepdims=True)
    Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    R = np.clip(Z @ Z.T, -1, 1)
    np.fill_diagonal(R, np.nan)
    s = c = 0.0
    for bi in shard_tile_rows(N, rank, world):     # what sie_corr_tau(shard_rank, shard_count) covers: rows of the upper triangle
        rows = slice(128 * bi, min(N, 128 * bi + 128))
        blk = R[rows, :]
        upper = np.triu(np.ones_like(R, dtype=bool), 1)[rows, :]
        m = upper & (blk >= 0) & (blk > rc)
        s += 2.0 * blk[m].sum()
        c += 2 * int(m.sum())
    tau = tau_from_shards(torch.tensor([s], dtype=torch.float64), torch.tensor([c], dtype=torch.int64))
    if rank == 0:
        full = R[(R >= 0) & (R > rc)]
        q.put(bool(abs(tau.item() - full.mean()) <= 1e-12 * abs(full.mean())))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_tau_allreduce_two_ranks():
    """the row-sharded correlation build exchanges only (sum, count); two gloo ranks reproduce the
    unsharded tau of ComplexNetworks.py:41-47."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tau_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_partition_sst_permutes_problems_and_meta_together():
    """RetrospectiveSweep launches each wave's GP twice (problems that read no SST network / those that do):
    SweepPlan.partition_sst reorders the records inside the wave ranges; the result lookup through prob_meta must not
    notice."""
    from seaiceextentforecasting_b200.forecast import GP_RESULT_DTYPE
    p = SweepPlan(NORTH_INITS, _sie(), 1985, 2020)
    q = SweepPlan(NORTH_INITS, _sie(), 1985, 2020)
    T = q.job_T[q.prob["job_sic"]]
    cut = int((T > 12).sum())
    ranges = [(0, cut), (cut, q.P)]
    splits = q.partition_sst(ranges)
    assert q.P == p.P == 432 and sorted(q.prob_meta) == sorted(p.prob_meta)
    for (p0, p1), sp in zip(ranges, splits):
        assert p0 < sp < p1
        assert (q.prob["job_sst"][p0:sp] < 0).all() and (q.prob["job_sst"][sp:p1] >= 0).all()
        assert (np.diff(q.job_T[q.prob["job_sic"][p0:sp]]) <= 0).all()        # stable: still longest windows first
    where = {m: i for i, m in enumerate(p.prob_meta)}
    for i, m in enumerate(q.prob_meta):                                       # record i of q is record where[m] of p
        assert q.prob[i] == p.prob[where[m]]
    # results written in either order assemble to the same dict
    rng = np.random.default_rng(0)
    raw_p = np.zeros(p.P, dtype=GP_RESULT_DTYPE)
    raw_p["fmean"], raw_p["fvar"] = rng.normal(size=p.P), rng.uniform(0.1, 1.0, size=p.P)
    raw_q = raw_p[[where[m] for m in q.prob_meta]]
    a, b = p.assemble(raw_p), q.assemble(raw_q)
    assert a.keys() == b.keys()
    for cfg in a:
        for k in a[cfg]:
            assert np.array_equal(np.asarray(a[cfg][k]), np.asarray(b[cfg][k]), equal_nan=True), (cfg, k)


def test_bench_ensemble_members_are_perturbations_of_the_base():
    """bench.py: rank r runs ensemble member r = base + 0.05 N(0,1); member 0 is the base realisation, the
    land mask and the saturated samples never change, concentrations stay in [0, 1]."""
    import bench
    w0, w0b, w3 = bench.make_workload(0), bench.make_workload(0), bench.make_workload(3)
    for name in w0["sic"]:
        a, b = w0["sic"][name], w3["sic"][name]
        assert np.array_equal(a, w0b["sic"][name], equal_nan=True)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        sat = (a == 0.0) | (a == 1.0)
        assert np.array_equal(a[sat], b[sat])
        assert np.nanmin(b) >= 0.0 and np.nanmax(b) <= 1.0
        d = (b - a)[~np.isnan(a) & ~sat]
        assert 0.03 < d.std() < 0.06 and abs(d.mean()) < 1e-3
    assert np.array_equal(np.isnan(w0["sst"]), np.isnan(w3["sst"]))
    assert all(np.array_equal(w0["sie"][k], w3["sie"][k]) for k in w0["sie"])
