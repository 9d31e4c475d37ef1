#!/usr/bin/env python
"""bench.py -- retrospective forecasts/sec (network + GPR) on the north retrospective sweep (BASELINE.json
configs[1]: 1985-2020 x June/July/August/September inits x Pan-Arctic/Beaufort/Chukchi = 432 forecasts from
144 SIC (57x57) + 36 SST (26x90) network builds, synthetic NSIDC-shaped monthly fields 1979-2020).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]        # reference algorithm on the host cores

One step = one pass of the whole hot path (detrend -> correlation/tau -> domain growth/merge -> node series/links
-> batched GP) over one sweep.  N>1: one process per GPU (torchrun), each rank runs an independent perturbed-SIC
ensemble member of the same sweep (no data-path collective; weak scaling), value = all forecasts / max-over-ranks
device time.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FMIN, FMAX, FIRST = 1985, 2020, 1979
TFULL = FMAX - FIRST + 1
METRIC = "retrospective forecasts/sec (network+GPR)"
WORKLOAD = ("north retrospective sweep 1985-2020 x June/July/August/September inits x 3 regions: 432 forecasts, "
            "144 SIC 57x57 + 36 SST 26x90 network builds, T=7..42")


def make_workload(member=0):
    """Synthetic inputs of BASELINE.json configs[1].  Member 0 is the base realisation; member m > 0 is the perturbed-SIC
    ensemble member of SURVEY.md 8(d) / BASELINE.json configs[4]: base field + 0.05 * N(0,1) with seed base_seed + m
    (concentrations stay in [0, 1], saturated samples and land stay as they are; the SIE targets are shared)."""
    from seaiceextentforecasting_b200 import synthetic as syn
    from seaiceextentforecasting_b200.config import CONFIGS, NORTH_INITS
    sic = {}
    for i, name in enumerate(NORTH_INITS):
        sic[name], _ = syn.make_field(57, 57, TFULL, 1000 + i)
    sie = dict(zip(CONFIGS["north_june"].regions, syn.make_sie(sic["north_september"], TFULL, 1000)))
    sst, _ = syn.make_field(26, 90, TFULL, 2000, latlon=True, saturate=False, n_modes=60, noise=0.6,
                            blob=(2.0, 5.0))
    if member > 0:
        rng = np.random.default_rng(5000 + member)
        for name in NORTH_INITS:
            f = sic[name]
            inside = (f > 0.0) & (f < 1.0)              # NaN (land) compares False
            sic[name] = np.where(inside, np.clip(f + 0.05 * rng.standard_normal(f.shape), 0.0, 1.0), f)
        sst = sst + 0.05 * rng.standard_normal(sst.shape)      # NaN land stays NaN
    return dict(sic=sic, sie=sie, sst=sst, psar=syn.make_psar(57, 57), lat=syn.make_lat_grid(26, 90))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True):
        # gpu_index: an index or a comma-separated list.  One poller (rank 0, all local GPUs) per job: every nvidia-smi
        # query takes driver locks, so one poller per rank perturbs the launches of all ranks
        self.idx, self.proc, self.lines, self.enabled = gpu_index, None, [], enabled

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.enabled:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                mx.append(cmax)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # region shorter than the sampling period: fall back to all samples
            for t, line in self.lines:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_task(args):
    """One (config, year) task of the sweep with the oracle port (reference algorithm, numpy/scipy)."""
    import warnings
    warnings.simplefilter("ignore")
    from oracle import sweep as osweep
    from seaiceextentforecasting_b200.config import CONFIGS
    name, year, member = args
    w = _ref_task.cache.get(member)
    if w is None:
        w = make_workload(member)
        cfgs = [CONFIGS[n] for n in w["sic"]]
        w["tables"] = {reg: osweep.sie_tables(w["sie"][reg], FMIN, FMAX) for reg in cfgs[0].regions}
        _ref_task.cache[member] = w
    cfg = CONFIGS[name]
    sie_dt = {r: w["tables"][r][0] for r in cfg.regions}
    sie_tr = {r: w["tables"][r][1] for r in cfg.regions}
    t0 = time.perf_counter()
    try:
        res, _ = osweep.run_job(cfg, year, w["sic"][name], w["psar"], sie_dt, sie_tr, FMIN, w["sst"], w["lat"])
        n = len(res)
    except (ValueError, IndexError, np.linalg.LinAlgError):
        n = 3   # the reference raises when <2 predictors pass; the work up to that point was still done
    return n, time.perf_counter() - t0


_ref_task.cache = {}


def sample_tasks(count):
    """A bounded, T-representative sample of the 144 (init, year) tasks: years spread evenly over 1985-2020."""
    from seaiceextentforecasting_b200.config import NORTH_INITS
    years = np.linspace(FMIN, FMAX, count).round().astype(int)
    return [(NORTH_INITS[i % 4], int(y), 0) for i, y in enumerate(years)]


def cpu_baseline_single(n_tasks=4):
    tasks = sample_tasks(n_tasks)
    t0 = time.perf_counter()
    n = sum(_ref_task(t)[0] for t in tasks)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "forecasts/s", "cores": 1, "kind": "port",
            "sample": f"{len(tasks)} of 144 (init,year) tasks = {n} of 432 forecasts, years "
                      f"{[t[1] for t in tasks]}, {dt:.1f} s, oracle port (numpy/scipy, bitmap lookups instead of the "
                      "reference's list scans: ~9x faster than the literal reference on area_level)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 48))
    n_tasks = max(procs, 8)
    tasks = sample_tasks(n_tasks)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_task, tasks[:procs])
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            n += sum(r[0] for r in pool.map(_ref_task, tasks))
        dt = time.perf_counter() - t0
    value = n / dt
    line = {"metric": METRIC, "value": value, "unit": "forecasts/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "step": f"bounded sample: {n_tasks} of 144 (init,year) tasks per step"},
            "cpu_baseline": {"value": value, "unit": "forecasts/s", "cores": procs, "kind": "port",
                             "sample": f"{n_tasks} (init,year) tasks per step over a {procs}-process pool "
                                       f"({cores} host cores); oracle port of the reference algorithm"},
            "e2e": {"value": value, "unit": "forecasts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def fp64_gemm_peak(torch):
    """cuBLAS DGEMM burst peak (TFLOP/s): the denominator for the DMMA correlation kernel, measured here because
    MEASURED_PEAKS.json only carries HBM and bf16 figures."""
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def corr_25km(torch, fp64_peak, world, rank):
    """BASELINE.json's second figure: all-pairs correlation + tau on the 25 km 448x304 grid (configs[3]), R not stored,
    tile rows sharded over the ranks (`sie_corr_tau(shard_rank, shard_count)`); TFLOP/s of the upper-triangle
    algorithmic work N(N+1)T against the FP64 tensor peak measured above.  Timed on the device, rank 0's shard."""
    from seaiceextentforecasting_b200 import synthetic as syn
    from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
    X, Y, T = 448, 304, 42
    data, _ = syn.make_field(X, Y, T, 7, n_modes=200)
    n_upper = int((~np.isnan(data).any(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=8)
    fields = h2d(data.reshape(1, X * Y, T))
    jf = torch.zeros(1, dtype=torch.int32, device="cuda")
    jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
    eng.detrend_zscore(fields, jf, jT, True)
    for _ in range(3):
        eng.corr_tau(rc, store_R=False, shard_rank=rank, shard_count=world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        eng.corr_tau(rc, store_R=False, shard_rank=rank, shard_count=world)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    N = int(eng.n_nodes.item())
    flop = N * (N + 1.0) * T / world
    tf = flop / (ms * 1e-3) / 1e12
    del eng
    torch.cuda.empty_cache()
    return {"workload": f"448x304 grid, {N} nodes, T={T}, R not stored, row shard {rank}/{world}", "ms": ms,
            "tflops_fp64": tf, "peak_tflops_fp64": fp64_peak, "frac": tf / fp64_peak if fp64_peak else None,
            "flop_model": "N(N+1)T per network (upper triangle), split evenly over the shards"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as entry
    from seaiceextentforecasting_b200 import build as b
    if rank == 0 and b.needs_build():
        entry.build()
    if world > 1:
        dist.barrier()
    from seaiceextentforecasting_b200.config import NORTH_INITS
    from seaiceextentforecasting_b200.forecast import RetrospectiveSweep

    w = make_workload(member=rank)
    sw = RetrospectiveSweep(NORTH_INITS, w["sic"], w["sie"], FMIN, FMAX, w["psar"], w["sst"], w["lat"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident steps (inputs already in HBM)
    sw.upload()
    for _ in range(args.warmup):
        sw.compute()
    sync_all()
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local, enabled=(rank == 0))
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        sw.compute()
    e1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1)
    # per-stage device times (for the roofline block): same work, one batch per grid so that every kernel is
    # bracketed by events on its own stream; not part of the headline timing
    marks_all = []
    for _ in range(args.steps):
        marks = []
        sw.compute(marks, waves=1)
        marks_all.append(marks)
    sync_all()
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    nf = torch.tensor([float(sw.n_forecasts)], dtype=torch.float64, device="cuda")
    per_rank_ms = [dev_ms / args.steps]
    if world > 1:
        allms = [torch.zeros_like(tmax) for _ in range(world)]
        dist.all_gather(allms, tmax)
        per_rank_ms = [float(t.item()) / args.steps for t in allms]     # rank r runs ensemble member r (different fields)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(nf, op=dist.ReduceOp.SUM)
    ms_per_step = float(tmax.item()) / args.steps
    total_forecasts = float(nf.item())
    value = total_forecasts / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API: pinned host -> device, compute, device -> host
    for _ in range(2):
        sw.run()
    sync_all()
    t0 = time.perf_counter()
    for out in sw.run_many(args.steps):      # every step: H2D of its inputs, the hot path, D2H of its results; the host
        pass                                 # assembles step i while step i+1 is on the device (forecast.run_many)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = total_forecasts / (float(e2e_t.item()) / args.steps)
    bad = int((sw.raw["info"] != 0).sum())
    sw.check_status()

    # ---------------- per-stage device time -> dominant kernel and its roofline
    stage_ms = {}
    for marks in marks_all:
        last = {}
        for name, ev in marks:          # consecutive marks of one chain (sic.* / sst.* / gp.*) bracket a stage
            tag = name.split(".")[0]
            if tag in last and not name.endswith(".start"):
                stage_ms[name] = stage_ms.get(name, 0.0) + last[tag].elapsed_time(ev)
            last[tag] = ev
    stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
    N = sw.sic.n_nodes.cpu().numpy().astype(np.int64)
    T = sw.plan.job_T.astype(np.int64)
    Ns = sw.sst.n_nodes.cpu().numpy().astype(np.int64)
    Ts = sw.plan.sst_T.astype(np.int64)
    corr_flop = float((N * (N + 1) * T).sum())
    corr_bytes = float((8 * N * N).sum())
    area_work = float(sw.sic.area_work.cpu().numpy()[:, 0].sum())
    cells = 57 * 57
    detr_bytes = float((16 * cells * T).sum())
    kern = {
        "sic.detrend_zscore": {"bound": "hbm", "work": detr_bytes + float((8 * N * sw.sic.Tp).sum())},
        # R is materialised here (8 N^2 bytes per network against N(N+1)T flop, T <= 42 -> <= 5.3 flop/B): HBM-store bound
        "sic.corr_tau": {"bound": "hbm", "work": corr_bytes + float((8 * N * sw.sic.Tp).sum()), "flop": corr_flop},
        "sic.area_level": {"bound": "hbm", "work": 8.0 * area_work},
    }
    top = max((k for k in stage_ms), key=lambda k: stage_ms[k])
    top = {"gp.gp": "gp"}.get(top, top)
    fp64_peak = fp64_gemm_peak(torch) if rank == 0 else 0.0
    roof = {}
    for name, k in kern.items():
        ms = stage_ms.get(name)
        if not ms:
            continue
        if k["bound"] == "hbm":
            ach = k["work"] / (ms * 1e-3) / 1e9
            roof[name] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                          "traffic": None, "ms": ms, "peak_source": hbm_src}
            if "flop" in k:
                roof[name]["tflops_fp64"] = k["flop"] / (ms * 1e-3) / 1e12
        else:
            ach = k["work"] / (ms * 1e-3) / 1e12
            roof[name] = {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": ach / fp64_peak if fp64_peak else None, "traffic": None, "ms": ms,
                          "peak_source": "cuBLAS DGEMM 6144^3 burst measured in this run (MEASURED_PEAKS.json has "
                                         "no FP64 figure; DMMA and DFMA peaks are nominally equal on B200)",
                          "store_GBps": k["bytes"] / (ms * 1e-3) / 1e9}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for name in roof:
            if name in traffic:
                roof[name]["traffic"] = traffic[name]
    except (OSError, ValueError):
        pass
    main_roof = dict(roof.get(top, roof.get("sic.area_level", {})))
    main_roof["kernel"] = top
    corr25 = corr_25km(torch, fp64_peak, world, rank) if args.corr25 else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "forecasts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "grid": "57x57 SIC + 26x90 SST", "years": [FMIN, FMAX],
                       "ensemble_members": world, "parallelism": f"task-parallel x{world} (one member per GPU)",
                       "members": "rank 0 = base realisation, rank r = base + 0.05*N(0,1) (perturbed-SIC ensemble member r)",
                       "l2": "working set per step ~7.8 GB (R matrices) >> 126 MB L2, no flush needed",
                       "schedule": (f"{len(sw.waves)} waves on separate streams, window-length edges T={list(sw.wave_T)}: "
                                    + "; ".join(f"{jr[1]-jr[0]} SIC networks + {pr[1]-pr[0]} GP problems" for (jr, sr, pr) in sw.waves)
                                    + ("; each wave's GP = a launch for the SIC-only problems + one for the SST-reading ones" if sw.gp_sst else "")
                                    + "; stage_ms / roofline timed in a separate single-wave pass; e2e loop = RetrospectiveSweep.run_many"
                                      " (step i's D2H + host assemble overlap step i+1)") if sw.multi_wave else "single wave"},
            "e2e": {"value": e2e_value, "unit": "forecasts/s", "h2d_bytes_per_step": sw.h2d_bytes(),
                    "d2h_bytes_per_step": sw.d2h_bytes(), "ms_per_step": 1e3 * float(e2e_t.item()) / args.steps},
            "gpu_launches": args.steps * sw.kernel_launches(),
            "clocks": clocks, "per_rank_ms_per_step": per_rank_ms,
            "roofline": main_roof,
            "roofline_all": roof,
            "corr_25km": corr25,
            "stage_ms": stage_ms,
            "gp_failures": bad,
            "checks": {"forecasts_per_rank": sw.n_forecasts, "areas_mean": float(sw.sic.n_areas.cpu().numpy().mean())},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-corr25", dest="corr25", action="store_false",
                    help="skip the 25 km all-pairs correlation probe (second half of BASELINE.json's metric)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set) are pointed at stderr, and the line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
