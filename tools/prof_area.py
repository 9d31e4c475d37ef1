"""Domain-growth kernel variants side by side on the SIC networks of the bench sweep (M members in one batch):
max_areas = 768 selects <256 threads, 16-bit indices, 2 CTAs/SM>, max_areas = 1024 selects <512 threads, 16-bit, 1 CTA/SM>
(csrc/area.cu plan_area).  Prints the launch time and, with a SIE_AREA_TIMERS=1 build, the per-phase SM cycles summed
over the jobs.   python tools/prof_area.py [members]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ws = [bench.make_workload(m) for m in range(M)]
names = ['seed', 'argmax', 'create', 'update+gather', 's2.select', 's2.discover', 's2.lists', 's2.rowmeans', 's2.stat',
         's2.fill', 'eval']
for MA in (768,) if len(sys.argv) > 2 else (768, 1024):
    sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                            [w["sst"] for w in ws], ws[0]["lat"], max_areas=MA)
    sw.upload()
    sw.compute(waves=1)
    torch.cuda.synchronize()
    for tag, eng in (("sic", sw.sic), ("sst", sw.sst)):
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.area_level()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        wk = eng.area_work.cpu().numpy()
        tot = (wk[:, 1] + wk[:, 2]).astype(np.float64)
        print(f"M={M} MA={MA} {tag}: {eng.B} jobs, area_level {best:.2f} ms ({best / M:.2f} per member); job cycles sum "
              f"{tot.sum() / 1e6:.0f} M (step1 {wk[:, 1].sum() / 1e6:.0f}, step2 {wk[:, 2].sum() / 1e6:.0f}), "
              f"mean {tot.mean() / 1e6:.2f} max {tot.max() / 1e6:.2f} M; steps {int((wk[:, 3] >> 32).sum())} rounds "
              f"{int((wk[:, 3] & 0xffffffff).sum())} slow {int(wk[:, 15].sum())}")
        if wk[:, 4:15].sum() > 0:
            print("   phases Mcyc:", {n: round(float(wk[:, 4 + i].sum()) / 1e6, 1) for i, n in enumerate(names)},
                  "merge", round(float(wk[:, 25].sum()) / 1e6, 1), "owner-wait", round(float(wk[:, 24].sum()) / 1e6, 1))
            print("   BK warp Mcyc outside/detect/assign/init/tail:",
                  [round(float(wk[:, 16 + i].sum()) / 1e6, 1) for i in range(5)])
            rb = wk[:, 26:32].astype(np.uint64)
            cnt, cyc = (rb >> np.uint64(44)).sum(axis=0), (rb & np.uint64((1 << 44) - 1)).sum(axis=0)
            print("   merge rounds by best-area size (<=4, <=8, <=16, <=32, <=64, >64): count", cnt.tolist(),
                  "Mcyc", [round(float(c) / 1e6, 1) for c in cyc], "kcyc/round", [round(float(c) / max(1, int(n)) / 1e3, 1) for c, n in zip(cyc, cnt)])
    del sw
    torch.cuda.empty_cache()
