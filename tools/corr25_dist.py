"""BASELINE configs[3] across GPUs: the 25 km 448x304 all-pairs correlation, row-sharded over the ranks of one box.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/corr25_dist.py
Every rank holds Z (N x T, 22 MB, recomputed locally from the same field), computes the 128-row tile rows
`bi % world == rank` of the upper triangle (R never stored, never crosses NVLink) and the ranks all-reduce 16 bytes
(sum, count) over NCCL to get the global tau.  Timing: barrier + CUDA events around K2 + the all-reduce, MAX over ranks."""
import sys, os, json; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
from seaiceextentforecasting_b200.parallel import tau_from_shards

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, Y, T = 448, 304, 42
data, _ = syn.make_field(X, Y, T, 7)
C = X * Y
n_upper = int((~np.isnan(data).any(axis=2)).sum())
eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=8)
fields = h2d(data.reshape(1, C, T))
jf = torch.zeros(1, dtype=torch.int32, device="cuda"); jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
eng.detrend_zscore(fields, jf, jT, True)
torch.cuda.synchronize()
N = int(eng.n_nodes.item())


def step():
    eng.corr_tau(rc, store_R=False, shard_rank=rank, shard_count=world)
    return tau_from_shards(eng.tau_sum, eng.tau_cnt)


for _ in range(3):
    tau = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    tau = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    flop = N * (N + 1.0) * T
    print(json.dumps({"workload": "448x304 grid all-pairs correlation + tau, R not stored, 128-row tile rows round-robin over ranks",
                      "n_gpus": world, "nodes": N, "T": T, "ms_per_build": float(ms.item()), "tflops_fp64_aggregate": flop / float(ms.item()) / 1e9,
                      "tau": float(tau.item()), "collective": "2 x all_reduce of one scalar (NCCL)" if world > 1 else "none"}))
if world > 1:
    dist.destroy_process_group()
