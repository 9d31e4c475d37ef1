"""Step time of the 8-member north sweep under the schedule variants: python tools/wave_check.py [members]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ws = [bench.make_workload(m) for m in range(M)]
for wave_T in ((12,), (9,), (16,), (10, 20)):
    sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                            [w["sst"] for w in ws], ws[0]["lat"], wave_T=wave_T)
    sw.upload()
    for mode in ((None, 1) if wave_T == (12,) else (None,)):
        for _ in range(3):
            sw.compute(waves=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            sw.compute(waves=mode)
        e1.record()
        torch.cuda.synchronize()
        print("wave_T", wave_T, "waves", mode, "ms/step", round(e0.elapsed_time(e1) / 6, 2))
    del sw
    torch.cuda.empty_cache()
