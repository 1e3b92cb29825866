import sys, math; sys.path.insert(0,'/root/repo')
import numpy as np, multilevel_mc_sdes_b200 as mm, time
res={}
for samp,n in (("f32bm",int(2e11)),("f64bm",int(6e10))):
    ms=[];vs=[]
    for seed in range(4):
        t=time.time()
        s=mm.Euro_Merton(seed=200+seed,sampler=samp,verbose=False).looper(n,0,2,False)
        m=s[2]/n; v=s[3]/n-m*m; ms.append(m); vs.append(v)
        print(samp,seed,'mean %.7f var %.5f  (%.1fs)'%(m,v,time.time()-t),flush=True)
    res[samp]=(np.mean(ms),np.mean(vs),n*4)
a,b=res["f32bm"],res["f64bm"]
se=math.sqrt(a[1]/a[2]+b[1]/b[2])
print('diff %+.3e +- %.1e  z %+.2f   var ratio %.6f'%(a[0]-b[0],se,(a[0]-b[0])/se,a[1]/b[1]))
