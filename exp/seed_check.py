import sys; sys.path.insert(0,'/root/repo')
import numpy as np, multilevel_mc_sdes_b200 as mm
for l,M in ((0,4),(1,2),(1,4),(3,4)):
    r=[mm.Euro_GBM(seed=s, verbose=False).looper(4_000_000, l, M, False) for s in (0,1,2,3)]
    print(l,M,[float(x[2]/4e6) for x in r])
# large ids at l=0
for off in (0, 2**32, 2**34, 10**11):
    o=mm.Euro_GBM(seed=0, verbose=False); o._next_path[0]=off
    s=o.looper(40_000_000,0,4,False); print('off',off, s[2]/4e7, s[3]/4e7-(s[2]/4e7)**2)
o=mm.Euro_GBM(verbose=False); print('BS',o.BS())
