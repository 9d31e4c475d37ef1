"""BASELINE configs[3]: a retrospective sweep on the full-resolution 25 km 448x304 Arctic grid, end to end across the
GPUs of one box.   Launch (N = 1, 2, 4, 8):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/sweep25_dist.py [years]
`years` target years ending in 2020 (default: one per GPU), three regions, predictors selected like the regional
forecasts of north/September1st.py:178-181 ((r > 0) & (p/2 < alpha)) for every region: with ~700 domains at this
resolution the "all areas" / "r > 0" rules of the 100 km scripts would select more predictors than one GP CTA handles
(max_pred <= 520, csrc/gp.cu), which the sweep reports loudly instead of returning NaNs.  Every (year) task = one 25 km network build (K1-K6; ~63.6 k nodes, the 32 GB correlation matrix stays on the
GPU that builds it) + its three GP forecasts.  Tasks are split over the ranks with no data-path collective
(RetrospectiveSweep(rank, world)); afterwards the labels / node series of all networks are all-gathered over NCCL
(parallel.all_gather_networks: the north_star's "all-gather of domain labels and node series only") and the GP records
are all-gathered as tensors.  Rank 0 prints one JSON line; with `--check` every rank also re-runs rank 0's first task
alone and compares bit for bit (sharding must not change results)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from seaiceextentforecasting_b200 import parallel
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200.config import CONFIGS, RULE_POS_SIG, ForecastConfig
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_years = int(args[0]) if args else world
check = "--check" in sys.argv
X, Y, FMAX = 448, 304, 2020
FMIN = FMAX - n_years + 1
Tfull = FMAX - 1979 + 1
t0 = time.perf_counter()
field, _ = syn.make_field(X, Y, Tfull, 7, n_modes=200)
S9 = CONFIGS["north_september"]
CFG = ForecastConfig("north_25km", "north", S9.regions, S9.ell, S9.sig, (RULE_POS_SIG,) * 3, alpha=0.01)
sie = dict(zip(CFG.regions, syn.make_sie(field, Tfull, 7)))
psar = syn.make_psar(X, Y)
t_gen = time.perf_counter() - t0

sw = RetrospectiveSweep([CFG], {"north_25km": field}, sie, FMIN, FMAX, psar, max_areas=20000, max_pred=512,
                        rank=rank, world=world)
sw.upload()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
sw.compute()
e1.record()
state = parallel.all_gather_networks(parallel.network_state(sw.sic), n_years)          # labels + node series, NCCL
recs = parallel.gather_records(sw.plan, None, device="cuda", raw_dev=sw.gp.out)      # GP records, NCCL
e2.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e2)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
sw.check_status(sw.download())
out = sw.plan.assemble(*recs)
gathered_bytes = sum(int(v.numel() * v.element_size()) for v in state.values())
ok = None
if check:
    one = RetrospectiveSweep([CFG], {"north_25km": field}, sie, FMIN, FMAX, psar, max_areas=20000, max_pred=512,
                             rank=0, world=n_years)                       # = the single task (year FMAX) alone
    ref = one.run()
    key = "Pan-Arctic_raw_fmean"
    i = FMAX - FMIN
    same_gp = all(np.array_equal(np.asarray(ref["north_25km"][r + s])[i:i + 1], np.asarray(out["north_25km"][r + s])[i:i + 1])
                  for r in CFG.regions for s in ("_raw_fmean", "_raw_fvar"))
    nA = int(one.sic.n_areas[0].item())
    same_net = (int(state["n_areas"][0].item()) == nA and torch.equal(state["label"][0], one.sic.label[0])
                and torch.equal(state["anomaly"][0, :nA], one.sic.anomaly[0, :nA]))
    flag = torch.tensor([1.0 if (same_gp and same_net) else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item() == 1.0)
if rank == 0:
    raw_all = recs[0]
    print(json.dumps({
        "workload": f"BASELINE configs[3]: 25 km 448x304 retrospective sweep, target years {FMIN}-{FMAX} (T = {FMIN - 1979 + 1}.."
                    f"{Tfull}), significance-selected predictors, 3 regions: {n_years} network builds of {int(sw.sic.n_nodes[0].item())} nodes + "
                    f"{3 * n_years} forecasts over {world} GPU(s)",
        "n_gpus": world, "ms_compute_max_over_ranks": float(ms[0].item()), "ms_with_allgathers": float(ms[1].item()),
        "forecasts_per_s": 3 * n_years / (float(ms[1].item()) * 1e-3),
        "allgathered_label_and_series_bytes": gathered_bytes, "areas_per_network": [int(x) for x in state["n_areas"].cpu().numpy()],
        "predictors": [int(x) for x in raw_all["n_pred"]], "info": [int(x) for x in raw_all["info"]],
        "fmean_pan_arctic": [float(x) for x in out["north_25km"]["Pan-Arctic_raw_fmean"]],
        "every_rank_reproduces_single_task_bit_for_bit": ok, "host_field_generation_s": t_gen}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
