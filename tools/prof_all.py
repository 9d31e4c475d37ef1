"""Two north sweep steps (every kernel of the path launches twice per step and wave) -- the target of the ncu passes whose
summaries are committed under profiles/ (see profiles/r01_kernels_v9_ncu.md for the exact command lines)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w = bench.make_workload(0)
sw = RetrospectiveSweep(NORTH_INITS, w['sic'], w['sie'], bench.FMIN, bench.FMAX, w['psar'], w['sst'], w['lat'])
for _ in range(2):
    out = sw.run()
torch.cuda.synchronize()
print("ok", type(out).__name__)
