"""Per-pass stage times of the serial (waves=0) schedule: python tools/stage_check.py [members | south]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
if len(sys.argv) > 1 and sys.argv[1] == "south":      # BASELINE configs[2]: south February 81x81 sweep, one member
    w = bench.make_workload_south(0)
    sw = RetrospectiveSweep(["south_february"], w["sic"], w["sie"], bench.FMIN, bench.FMAX, w["psar"], max_pred=512)
else:
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ws = [bench.make_workload(m) for m in range(M)]
    sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                            [w["sst"] for w in ws], ws[0]["lat"])
sw.upload()
for _ in range(3):
    sw.compute()
torch.cuda.synchronize()
for it in range(4):
    marks = []
    sw.compute(marks, waves=0)
    torch.cuda.synchronize()
    print(it, {k: round(v, 2) for k, v in bench.stage_times([marks], 1).items()})
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sw.sic.area_level(); e1.record(); torch.cuda.synchronize()
    print("sic.area_level alone", round(e0.elapsed_time(e1), 2))
