"""Multi-GPU build of one network whose correlation pass is row-sharded (BASELINE configs[3] structure).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/net_dist.py X Y T [n_modes]
Every rank recomputes the unit-norm rows z (replicated), computes the 128-row tile rows bi % N == rank of the tau-only
correlation pass, the (sum, count) partials are all-reduced over NCCL, rank 0 grows the domains from z (no stored
matrix) and builds the node series, and labels / area tables / node series are broadcast.  Every rank then checks its
copy against a single-GPU stored-matrix build it does itself (bit-exact domains, identical tau and node series)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from seaiceextentforecasting_b200 import parallel
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, Y, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nm = int(sys.argv[4]) if len(sys.argv) > 4 else None
data, _ = syn.make_field(X, Y, T, 7, n_modes=nm)
C = X * Y
n_upper = int((~np.isnan(data).any(axis=2)).sum())
MA = min(C // 2 + 1, 20000)
fields = h2d(data.reshape(1, C, T))
jf = torch.zeros(1, dtype=torch.int32, device="cuda")
jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
scale = h2d(np.sqrt(syn.make_psar(X, Y)).reshape(-1))

eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=MA)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
tau = parallel.build_networks_sharded(eng, fields, jf, jT, rc, scale)
e[1].record()
torch.cuda.synchronize()
ms = torch.tensor([e[0].elapsed_time(e[1])], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

# single-GPU reference build with the stored matrix on this rank (fits for the sizes this tool is run at)
ref = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=True, max_areas=MA)
ref.build(fields, jf, jT, rc, scale)
torch.cuda.synchronize()
nA = int(ref.n_areas.item())
same = (int(eng.n_areas.item()) == nA and torch.equal(eng.area_key[0, :nA], ref.area_key[0, :nA])
        and torch.equal(eng.area_start[0, :nA + 1], ref.area_start[0, :nA + 1])
        and torch.equal(eng.area_cells[0, :int(ref.area_start[0, nA])], ref.area_cells[0, :int(ref.area_start[0, nA])])
        and torch.equal(eng.label, ref.label) and torch.equal(eng.anomaly[0, :nA], ref.anomaly[0, :nA]))
tau_ok = abs(float(tau.item()) - float(ref.tau.item())) <= 1e-12 * abs(float(ref.tau.item()))
flag = torch.tensor([1.0 if (same and tau_ok) else 0.0], device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"workload": f"{X}x{Y}x{T} network, {int(eng.n_nodes.item())} nodes, correlation rows sharded over {world} rank(s), "
                                  "domains grown on rank 0 from z (no stored matrix), labels + node series broadcast",
                      "n_gpus": world, "ms_build_max_over_ranks": float(ms.item()), "areas": nA, "tau": float(tau.item()),
                      "every_rank_equals_single_gpu_stored_matrix_build": bool(flag.item() == 1.0)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
