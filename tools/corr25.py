"""Config 4 probe: all-pairs correlation + tau on the 25 km 448x304 grid (R not stored), one GPU, optional row shards."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
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
N = int(eng.n_nodes.item()); print("nodes", N, "ldn", eng.ldn)
for shards in (1, 2, 8):
    for _ in range(2):
        eng.corr_tau(rc, store_R=False, shard_rank=0, shard_count=shards)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.corr_tau(rc, store_R=False, shard_rank=0, shard_count=shards); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = N * (N + 1.0) * T / shards
    print(f"shards={shards} rank0: {ms:.2f} ms  {flop/ms/1e9:.2f} TFLOP/s (algorithmic N(N+1)T/shards)  tau_sum={eng.tau_sum.item():.6f} cnt={eng.tau_cnt.item()}")
# sharded partials add up to the unsharded result
eng.corr_tau(rc, store_R=False); torch.cuda.synchronize(); s1, c1 = eng.tau_sum.item(), eng.tau_cnt.item()
ss, cc = 0.0, 0
for r in range(4):
    eng.corr_tau(rc, store_R=False, shard_rank=r, shard_count=4); torch.cuda.synchronize(); ss += eng.tau_sum.item(); cc += eng.tau_cnt.item()
print("unsharded", s1, c1, "sum of 4 shards", ss, cc, "rel diff", abs(ss - s1) / abs(s1))
