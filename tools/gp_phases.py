"""Where k_gp_forecast's SM cycles go, per phase, summed over the GP problems of the 8-member north sweep (or the south
sweep).  Needs a library built with the phase counters:
    SIE_DEFINES=SIE_GP_TIMERS python seaiceextentforecasting_b200/build.py --force && python tools/gp_phases.py [members|south]"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200 import _lib
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep

NAMES = ["selection + design + Laplacian", "A^2 A^4 A^6 + norms", "order selection (A^8, A^10, power iterations <= 19)",
         "|A|^27 power iteration", "Pade U / V", "LU solve", "squarings", "X E, W (+ gradient operands)",
         "Cholesky x2, solves, predict"]
if len(sys.argv) > 1 and sys.argv[1] == "south":
    w = bench.make_workload_south(0)
    sw = RetrospectiveSweep(["south_february"], w["sic"], w["sie"], bench.FMIN, bench.FMAX, w["psar"], max_pred=512)
else:
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ws = [bench.make_workload(m) for m in range(M)]
    sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                            [w["sst"] for w in ws], ws[0]["lat"])
sw.upload()
sw.compute(waves=0)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 16)()
lib.sie_debug_gp_phases(buf)                    # clear
sw.compute(waves=0)
torch.cuda.synchronize()
assert lib.sie_debug_gp_phases(buf) == 0
cyc = np.array(list(buf), dtype=np.float64)[:len(NAMES)]
raw = sw.download()
tot = float(raw["cycles_total"].sum())
print(f"{sw.P} problems, {tot / 1e6:.0f} Mcycles in total (sum over problems), phases account for {cyc.sum() / 1e6:.0f}")
if cyc.sum() == 0:
    print("all counters zero: build with SIE_DEFINES=SIE_GP_TIMERS")
for n, c in zip(NAMES, cyc):
    print(f"  {c / 1e6:8.0f} Mcyc {100 * c / max(cyc.sum(), 1):5.1f} %  {n}")
