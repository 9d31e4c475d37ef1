"""configs[4] probe: the north sweep's 432 GP problems x the 20 x 20 hyper-parameter grid on one GPU."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w = bench.make_workload(0)
sw = RetrospectiveSweep(NORTH_INITS, w['sic'], w['sie'], bench.FMIN, bench.FMAX, w['psar'], w['sst'], w['lat'])
sw.run()
g = sw.hyper_grid()
torch.cuda.synchronize()
t0 = time.perf_counter(); g = sw.hyper_grid(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
ok = (g['info'] == 0)
print(f"{g.size} (problem, l, sigma) evaluations, {sw.P * g.shape[1]} expm, {dt*1e3:.1f} ms -> {g.size/dt:.0f} evaluations/s; "
      f"SPD {ok.mean()*100:.1f} %; Np max {g['n_pred'].max()}")
