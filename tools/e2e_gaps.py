"""Device timeline of the end-to-end loop (RetrospectiveSweep.run_many): per step the device time of compute() and the idle
gap before it.   python tools/e2e_gaps.py [members] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ws = [bench.make_workload(m) for m in range(M)]
sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                        [w["sst"] for w in ws], ws[0]["lat"])
for _ in range(2):
    sw.run()
marks = []
orig = sw.compute
def timed_compute(*a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    h0 = time.perf_counter()
    orig(*a, **k)
    h1 = time.perf_counter()
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    marks.append((e0, e1, h1 - h0))
sw.compute = timed_compute
torch.cuda.synchronize()
t0 = time.perf_counter()
th = []
for out in sw.run_many(steps):
    th.append(time.perf_counter())
torch.cuda.synchronize()
t1 = time.perf_counter()
print("e2e ms/step", round((t1 - t0) / steps * 1e3, 2))
print("device ms per compute", [round(a.elapsed_time(b), 2) for a, b, _ in marks])
print("idle gap before compute i (ms)", [round(marks[i][1].elapsed_time(marks[i + 1][0]), 2) for i in range(len(marks) - 1)])
print("host ms to enqueue a step", [round(h * 1e3, 2) for _, _, h in marks])
print("host ms between yields", [round((b - a) * 1e3, 2) for a, b in zip(th[:-1], th[1:])])
