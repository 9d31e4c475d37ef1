import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w=bench.make_workload(0)
sw=RetrospectiveSweep(NORTH_INITS,w['sic'],w['sie'],bench.FMIN,bench.FMAX,w['psar'],w['sst'],w['lat'])
sw.sic.area_scratch.zero_()
out=sw.run(); torch.cuda.synchronize()
eng=sw.sic
stride=eng.area_scratch_bytes//eng.B
C=eng.C
raw=eng.area_scratch.view(torch.int32).cpu().numpy()
for b in (0,143):
    off=(b*stride + (C*4+255)//256*256)//4
    d=raw[off:off+4*400].reshape(-1,4)
    d=d[d[:,0]>0]
    print('job',b,'rounds',len(d),'sum units',d[:,2].sum(),'sum kcyc',d[:,3].sum()*16/1e3)
    for r in d[:60]: print('  nb=%d nn=%d units=%d kcyc=%.1f'%(r[0],r[1],r[2],r[3]*16/1e3))
