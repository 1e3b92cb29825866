"""Compare level sums of two builds of the library on the same (seed, ids): exp/cmp_libs.py old.so new.so"""
import ctypes as C, os, subprocess, sys, json
if len(sys.argv) == 3:
    outs = []
    for lib in sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, "MLMCB_LIB": os.path.abspath(lib)}, capture_output=True, text=True)
        if r.returncode: print(r.stderr); sys.exit(1)
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    bad = 0
    for k in outs[0]:
        a, b = outs[0][k], outs[1][k]
        same = a == b
        bad += not same
        if not same: print("DIFF", k, a, b)
    print(f"{len(outs[0])} cases, {bad} differ")
    sys.exit(bad != 0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm
res = {}
for cls in ("Euro_Merton", "Asian_Merton", "Lookback_Merton", "Digital_Merton"):
    for M, ls in ((2, (0, 1, 2, 3, 5, 8)), (4, (1, 2, 4)), (3, (2,))):
        for l in ls:
            for kw in ({}, {"lam": 4.0}):
                o = getattr(mm, cls)(seed=11, **kw)
                res[f"{cls} M{M} l{l} {kw}"] = o.looper(100_003, l, M, False).tolist()
o = mm.Euro_Merton(seed=5, sampler="f64bm"); res["f64bm"] = o.looper(50_001, 3, 4, False).tolist()
o = mm.Euro_Merton(seed=5, scheme="milstein"); res["mil"] = o.looper(50_001, 3, 4, False).tolist()
print(json.dumps(res))
