import sys; sys.path.insert(0,'/root/repo')
import numpy as np, multilevel_mc_sdes_b200 as mm
bs = mm.Euro_GBM(verbose=False).BS()
rows=[]
for seed in (0,1,2,3,10,11,12,13):
    o = mm.Euro_GBM(verbose=False, seed=seed, alpha_0=1, beta_0=1)
    sums, N = o.mlmc(5e-5, M=4)
    m = sums[0]/N
    v = sums[1]/N - m*m
    rows.append(m)
    print(seed, 'err %+.3e'%(m.sum()-bs), 'L', len(N)-1, 'pred std %.2e'%np.sqrt((v/N).sum()), 'N0 %.3g'%N[0], 'means', ' '.join('%.7f'%x for x in m))
rows=np.array(rows)
print('per-level std across seeds', rows.std(axis=0))
