"""(needs a library built with per-phase timers: SIE_AREA_TIMERS=1 python seaiceextentforecasting_b200/build.py --force)
Phase cycles of k_area_level for ONE 57x57x42 network (B=1: no other CTA competes for the memory system)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w = bench.make_workload(0)
nj = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sw = RetrospectiveSweep(NORTH_INITS, w['sic'], w['sie'], bench.FMIN, bench.FMAX, w['psar'], w['sst'], w['lat'])
sw.upload(); sw.compute(waves=1); torch.cuda.synchronize()
eng = sw.sic
for rep in range(2):
    eng.area_level(jr=(0, nj)); torch.cuda.synchronize()
wk = eng.area_work.cpu().numpy()
names = ['seed','argmax','create','update+gather','s2.select','s2.discover','s2.lists','s2.rowmeans','s2.stat','s2.fill','eval']
print('jobs', nj, 'job0 step1 Mcyc', wk[0,1]/1e6, 'step2', wk[0,2]/1e6, {n: round(wk[0,4+i]/1e6,2) for i,n in enumerate(names)}, 'merge', round(wk[0,25]/1e6,2))
