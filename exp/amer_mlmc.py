import time, sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import multilevel_mc_sdes_b200 as mm
mm.Amer_GBM(verbose=False).looper(1000,2,2,False)
for eps in (1e-2, 2e-3):
    o = mm.Amer_GBM(verbose=False)
    torch.cuda.synchronize(); t=time.perf_counter()
    sums, N = o.mlmc(eps, M=2)
    dt=time.perf_counter()-t
    print(eps, round(dt,3), 'price', float(np.sum(sums[0]/N)), ['%.3g'%n for n in N])
