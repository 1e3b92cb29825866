"""f32bm vs f64bm Merton level means at large N (the two samplers build the jump records independently:
fp32 MUFU gap + fp32 Box-Muller vs fp64 log + fp64 Box-Muller)."""
import sys, json, math; sys.path.insert(0,'/root/repo')
import numpy as np, multilevel_mc_sdes_b200 as mm
rep={}
for cname in ("Euro_Merton","Asian_Merton","Digital_Merton"):
    for l,M,n in ((0,2,int(4e10)),(1,2,int(2e10)),(3,2,int(1e10)),(5,2,int(4e9))):
        r={}
        for samp in ("f32bm","f64bm"):
            nn = n if samp=="f32bm" else n//4
            s = getattr(mm,cname)(seed=55, sampler=samp, verbose=False).looper(nn,l,M,False)
            md, mf = s[0]/nn, s[2]/nn
            r[samp]=(md, s[1]/nn-md*md, mf, s[3]/nn-mf*mf, nn)
        a,b=r["f32bm"],r["f64bm"]
        zd=(a[0]-b[0])/math.sqrt(a[1]/a[4]+b[1]/b[4]); zf=(a[2]-b[2])/math.sqrt(a[3]/a[4]+b[3]/b[4])
        rep[f"{cname}_l{l}"]={"z_dP":zd,"z_Pf":zf,"dPf":a[2]-b[2],"se_Pf":math.sqrt(a[3]/a[4]+b[3]/b[4]),"var_ratio_Pf":a[3]/b[3],"var_ratio_dP":a[1]/b[1]}
        print(cname,l,"z(dP) %+.2f z(Pf) %+.2f  dPf %+.2e +- %.1e  VPf ratio %.6f VdP ratio %.6f"%(zd,zf,a[2]-b[2],rep[f"{cname}_l{l}"]["se_Pf"],a[3]/b[3],a[1]/b[1]), flush=True)
print(json.dumps(rep))
