"""Where the GP stage's SM cycles go: the problems of one north sweep bucketed by predictor count (kernel clock64 counters)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w = bench.make_workload(0)
sw = RetrospectiveSweep(NORTH_INITS, w['sic'], w['sie'], bench.FMIN, bench.FMAX, w['psar'], w['sst'], w['lat'])
sw.run(); sw.run()
raw = sw.raw
ok = raw["info"] == 0
Np, m, s = raw["n_pred"], raw["expm_m"], raw["expm_s"]
tot, ex = raw["cycles_total"] / 1e6, raw["cycles_expm"] / 1e6
print(f"{ok.sum()} problems, total {tot[ok].sum():.1f} Mcyc, expm {ex[ok].sum():.1f} Mcyc ({100 * ex[ok].sum() / tot[ok].sum():.0f} %)")
edges = [0, 16, 32, 48, 64, 96, 128, 192, 512]
for lo, hi in zip(edges[:-1], edges[1:]):
    sel = ok & (Np > lo) & (Np <= hi)
    if not sel.any():
        continue
    gem = np.where(m[sel] <= 5, 4, np.where(m[sel] <= 9, 5, 8)) + s[sel]
    tiles = ((Np[sel] + 63) // 64) ** 2 * ((Np[sel] + 15) // 16)
    steps = gem * tiles
    print(f"Np ({lo:3d},{hi:3d}]: {sel.sum():4d} problems, total {tot[sel].sum():7.1f} Mcyc ({100 * tot[sel].sum() / tot[ok].sum():4.1f} %), "
          f"expm {ex[sel].sum():7.1f}, mean s {s[sel].mean():5.1f}, m13 {np.mean(m[sel] == 13):.2f}, GEMM steps/problem {steps.mean():7.1f}, "
          f"expm cycles/GEMM step {1e6 * ex[sel].sum() / steps.sum():7.0f}, non-expm/problem {1e3 * (tot[sel] - ex[sel]).mean():6.0f} kcyc")
print("s histogram", np.bincount(s[ok]))
