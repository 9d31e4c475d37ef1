# needs a library built with per-phase timers: SIE_AREA_TIMERS=1 python seaiceextentforecasting_b200/build.py --force
import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from seaiceextentforecasting_b200.config import NORTH_INITS
from seaiceextentforecasting_b200.forecast import RetrospectiveSweep
w=bench.make_workload(0)
sw=RetrospectiveSweep(NORTH_INITS,w['sic'],w['sie'],bench.FMIN,bench.FMAX,w['psar'],w['sst'],w['lat'])
out=sw.run(); out=sw.run()
torch.cuda.synchronize()
for tag,eng,T in (('sic',sw.sic,sw.plan.job_T),('sst',sw.sst,sw.plan.sst_T)):
    wk=eng.area_work.cpu().numpy(); na=eng.n_areas.cpu().numpy(); nn=eng.n_nodes.cpu().numpy()
    order=np.argsort(-(wk[:,1]+wk[:,2]))
    print(tag,'job T N nA gathers step1_Mcyc step2_Mcyc steps rounds')
    for b in list(order[:8])+list(order[-3:]):
        print(tag,b,T[b],nn[b],na[b],wk[b,0],round(wk[b,1]/1e6,2),round(wk[b,2]/1e6,2),wk[b,3]>>32,wk[b,3]&0xffffffff)
    print(tag,'sum step1 Mcyc',wk[:,1].sum()/1e6,'sum step2',wk[:,2].sum()/1e6)
    names=['seed','argmax','create','update+gather','s2.select','s2.discover','s2.lists','s2.rowmeans','s2.stat','s2.fill','eval','slow-steps']
    for b in list(order[:3])+list(order[-2:]):
        print(tag,b,'phases Mcyc',{n:round(wk[b,4+i]/1e6,2) for i,n in enumerate(names)})
        print(tag,b,'BK Mcyc outside/detect/assign/init/tail',[round(wk[b,16+i]/1e6,2) for i in range(5)],'inits',wk[b,21],'owner-wait',round(wk[b,24]/1e6,2),'s2.merge',round(wk[b,25]/1e6,2),'slow',wk[b,15])
raw=sw.raw
order=np.argsort(-raw['cycles_total'])
print('gp: idx meta n npred m s info total_Mcyc expm_Mcyc')
for i in list(order[:12])+list(order[-3:]):
    print(i,sw.plan.prob_meta[i],sw.plan.prob[i]['n'],raw[i]['n_pred'],raw[i]['expm_m'],raw[i]['expm_s'],raw[i]['info'],round(raw[i]['cycles_total']/1e6,3),round(raw[i]['cycles_expm']/1e6,3))
print('gp sum total Mcyc',raw['cycles_total'].sum()/1e6,'expm',raw['cycles_expm'].sum()/1e6, 'bad',[(i,sw.plan.prob_meta[i],raw[i]['info'],raw[i]['n_pred']) for i in np.nonzero(raw['info'])[0]])
print('s hist',np.bincount(raw['expm_s']), 'npred max',raw['n_pred'].max(), 'mean', raw['n_pred'].mean())
# GP problems of the long-window wave (T > 12): what the tail of the step looks like
pm = np.array(sw.plan.prob_meta)
Tp = sw.plan.job_T[sw.plan.prob["job_sic"]]
for lo, hi in ((13, 20), (21, 30), (31, 42)):
    sel = np.nonzero((Tp >= lo) & (Tp <= hi))[0]
    c = raw['cycles_total'][sel] / 1e6
    print(f"gp T {lo}-{hi}: {len(sel)} problems, sum {c.sum():.1f} Mcyc, max {c.max():.2f}, Np max {raw['n_pred'][sel].max()}, mean {raw['n_pred'][sel].mean():.1f}, s max {raw['expm_s'][sel].max()}")
