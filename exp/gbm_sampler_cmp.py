"""f32bm vs f64bm at the headline level (l=6, M=4): E[dP], V[dP], E[Pf]."""
import sys, math; sys.path.insert(0,'/root/repo')
import numpy as np, multilevel_mc_sdes_b200 as mm
for cname in ("Euro_GBM","Asian_GBM","Lookback_GBM","Digital_GBM"):
    r={}
    for samp,n in (("f32bm",int(3e9)),("f64bm",int(1e9))):
        s=getattr(mm,cname)(seed=77,sampler=samp,verbose=False).looper(n,6,4,False)
        md,mf=s[0]/n,s[2]/n
        r[samp]=(md,s[1]/n-md*md,mf,s[3]/n-mf*mf,n)
    a,b=r["f32bm"],r["f64bm"]
    zd=(a[0]-b[0])/math.sqrt(a[1]/a[4]+b[1]/b[4]); zf=(a[2]-b[2])/math.sqrt(a[3]/a[4]+b[3]/b[4])
    print(cname,"E[dP] %.3e / %.3e  z %+.2f | V[dP] %.4e / %.4e ratio %.5f | E[Pf] z %+.2f"%(a[0],b[0],zd,a[1],b[1],a[1]/b[1],zf),flush=True)
