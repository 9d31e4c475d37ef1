"""Config 4 probe: the whole network build (K1-K6) of ONE large polar grid on one GPU, once with R materialised and once
without (tau-only correlation pass, domain growth recomputing correlations from the unit-norm rows); the two builds must
give identical domains.   usage: net25.py X Y T [n_modes] [max_areas]"""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
X, Y, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nm = int(sys.argv[4]) if len(sys.argv) > 4 else 200
data, _ = syn.make_field(X, Y, T, 7, n_modes=nm)
C = X * Y
n_upper = int((~np.isnan(data).any(axis=2)).sum())
MA = int(sys.argv[5]) if len(sys.argv) > 5 else None
fields = h2d(data.reshape(1, C, T))
jf = torch.zeros(1, dtype=torch.int32, device="cuda"); jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
scale = h2d(np.sqrt(syn.make_psar(X, Y)).reshape(-1))
names = ["detrend_zscore", "corr_tau", "area_level", "intra_links"]
tables = {}
for keep in (True, False):
    eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=keep, max_areas=MA)
    print("keep_R", keep, "cells", C, "nodes<=", n_upper, "ldn", eng.ldn, "R GB", eng.ldn ** 2 * 8 / 1e9 if keep else 0.0, "MA", eng.MA, flush=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); eng.detrend_zscore(fields, jf, jT, True)
    ev[1].record(); eng.corr_tau(rc, store_R=keep)
    ev[2].record(); eng.area_level()
    ev[3].record(); eng.intra_links(scale)
    ev[4].record(); torch.cuda.synchronize()
    print({n: round(ev[i].elapsed_time(ev[i + 1]), 2) for i, n in enumerate(names)}, "ms")
    wk = eng.area_work.cpu().numpy()[0]
    print("status", eng.status.item(), "areas", eng.n_areas.item(), "tau", eng.tau.item(), "growth steps", wk[3] >> 32, "merge rounds", wk[3] & 0xffffffff,
          "slow steps", wk[15], "gathers", wk[0])
    st = eng.area_start.cpu().numpy()[0][:eng.n_areas.item() + 1]
    sz = np.diff(st)
    print("largest areas", sorted(sz.tolist(), reverse=True)[:8], "cells in areas", int(sz.sum()))
    tables[keep] = (eng.n_areas.item(), eng.area_key.cpu().numpy().copy(), eng.area_start.cpu().numpy().copy(),
                    eng.area_cells.cpu().numpy().copy(), eng.label.cpu().numpy().copy(), eng.tau.item())
    del eng
    torch.cuda.empty_cache()
a, b = tables[True], tables[False]
nA = a[0]
same = (a[0] == b[0] and np.array_equal(a[1][0, :nA], b[1][0, :nA]) and np.array_equal(a[2][0, :nA + 1], b[2][0, :nA + 1])
        and np.array_equal(a[3][0, :a[2][0, nA]], b[3][0, :a[2][0, nA]]) and np.array_equal(a[4], b[4]))
print("domains identical with and without the stored matrix:", same, "tau rel diff", abs(a[5] - b[5]) / abs(a[5]))
assert same
