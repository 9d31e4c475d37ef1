"""Two steps of the batched north sweep (bench.py's step with `members` ensemble members): the target of the ncu passes
whose summaries are committed under profiles/ (r02_*).   python tools/prof_step.py [members] [steps] [waves]
waves: omitted = the product schedule; 0 = every stage alone on one stream (the launches bench.py's roofline pass times)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
waves = int(sys.argv[3]) if len(sys.argv) > 3 else None
ws = [bench.make_workload(m) for m in range(M)]
sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                        [w["sst"] for w in ws], ws[0]["lat"])
sw.upload()
for _ in range(steps):
    sw.compute(waves=waves)
torch.cuda.synchronize()
raw = sw.download()
print("ok", len(raw), "records,", int((raw["info"] == 0).sum()), "finite; kernels per step", sw.kernel_launches())
