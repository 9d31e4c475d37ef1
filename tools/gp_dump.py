"""Dump the per-problem GP records of one serial pass of the 8-member north sweep (cycles, Np, m, s, l, n, rule,
candidate count) for offline analysis of the LPT order: python tools/gp_dump.py out.npz [members]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ws = [bench.make_workload(m) for m in range(M)]
sw = RetrospectiveSweep(NORTH_INITS, [w["sic"] for w in ws], ws[0]["sie"], bench.FMIN, bench.FMAX, ws[0]["psar"],
                        [w["sst"] for w in ws], ws[0]["lat"])
sw.upload()
for _ in range(2):
    sw.compute(waves=0)
torch.cuda.synchronize()
raw = sw.download()
prob = sw.plan.prob
na_sic = sw.sic.n_areas.cpu().numpy(); na_sst = sw.sst.n_areas.cpu().numpy()
cand = na_sic[prob["job_sic"]] + np.where(prob["job_sst"] >= 0, na_sst[np.maximum(prob["job_sst"], 0)], 0)
np.savez_compressed(sys.argv[1], cycles=raw["cycles_total"], expm=raw["cycles_expm"], n_pred=raw["n_pred"], m=raw["expm_m"],
                    s=raw["expm_s"], info=raw["info"], ell=prob["ell"], n=prob["n"], rule=prob["rule"], cand=cand)
print("saved", len(raw))
