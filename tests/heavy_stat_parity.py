#!/usr/bin/env python
"""One-off heavy statistical parity run (north-star check (b) at small standard errors).

    python tests/heavy_stat_parity.py [f64bm|f32bm] > gpurun_out/heavy_stat_parity.json

For every product (8 + antithetic + Milstein + American put) at two levels: the CUDA level function
with N_gpu samples against the pinned C port of the reference (own generator, all host threads) with
N_cpu samples; z-scores of E[dP], E[Pf], Var[dP], Var[Pf] under the combined standard error
(variance s.e. from the sample fourth moments is not available from the 7 sums, so the variance
z-scores use the two estimates' spread under a normal-kurtosis bound and are reported as ratios).
Lives under tests/ because it uses the oracle."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402
from oracle import c_oracle, np_oracle as O  # noqa: E402

N_CPU, N_GPU = 40_000_000, 4_000_000_000
SAMPLER = sys.argv[1] if len(sys.argv) > 1 else "f64bm"
threads = os.cpu_count() or 1
cases = []
for cname, model, pay in (("Euro_GBM", O.GBM, O.EURO), ("Asian_GBM", O.GBM, O.ASIAN), ("Lookback_GBM", O.GBM, O.LOOKBACK),
                          ("Digital_GBM", O.GBM, O.DIGITAL), ("Euro_Merton", O.MERTON, O.EURO), ("Asian_Merton", O.MERTON, O.ASIAN),
                          ("Lookback_Merton", O.MERTON, O.LOOKBACK), ("Digital_Merton", O.MERTON, O.DIGITAL)):
    for l, M in ((3, 4), (5, 2)):
        cases.append((cname, model, pay, l, M, "em", False))
# the Merton walks: l=0 lane-queue loop, 2 steps event-driven (as are (3,4) and (5,2) above), 256 steps block-uniform
for cname, pay in (("Euro_Merton", O.EURO), ("Asian_Merton", O.ASIAN), ("Lookback_Merton", O.LOOKBACK), ("Digital_Merton", O.DIGITAL)):
    for l, M in ((0, 2), (1, 2), (8, 2)):
        cases.append((cname, O.MERTON, pay, l, M, "em", False))
cases += [("Euro_GBM", O.GBM, O.EURO, 3, 4, "milstein", False), ("Asian_Merton", O.MERTON, O.ASIAN, 3, 4, "milstein", False),
          ("Euro_GBM", O.GBM, O.EURO, 3, 4, "em", True), ("Asian_GBM", O.GBM, O.ASIAN, 4, 2, "em", True),
          ("Asian_GBM", O.GBM, O.ASIAN, 3, 4, "milstein", True),        # Milstein stepper under the antithetic estimator
          ("Euro_GBM", O.GBM, O.EURO, 1, 4, "em", False), ("Asian_GBM", O.GBM, O.ASIAN, 2, 2, "em", False),   # four-step kernels
          ("Amer_GBM", O.GBM, O.AMERICAN, 4, 2, "em", False), ("Amer_GBM", O.GBM, O.AMERICAN, 6, 2, "em", False)]
rep = {"N_cpu": N_CPU, "N_gpu": N_GPU, "cpu_threads": threads, "sampler": SAMPLER, "cases": []}
for cname, model, pay, l, M, scheme, anti in cases:
    p = O.Params(scheme=scheme)
    t0 = time.perf_counter()
    ref = c_oracle.level_sums_anti(p, pay, l, M, N_CPU, seed=77, nthreads=threads) if anti else \
        c_oracle.level_sums(p, model, pay, l, M, N_CPU, seed=77, nthreads=threads)
    t_cpu = time.perf_counter() - t0
    opt = getattr(mm, cname)(seed=4321, scheme=scheme, sampler=SAMPLER, verbose=False)
    t0 = time.perf_counter()
    s = opt.looper(N_GPU, l, M, anti)
    t_gpu = time.perf_counter() - t0

    def mom(x, n):
        md, mf = x[0] / n, x[2] / n
        return md, x[1] / n - md * md, mf, x[3] / n - mf * mf

    gd, gvd, gf, gvf = mom(s, N_GPU)
    cd, cvd, cf, cvf = mom(ref, N_CPU)
    z_d = (gd - cd) / np.sqrt(gvd / N_GPU + cvd / N_CPU)
    z_f = (gf - cf) / np.sqrt(gvf / N_GPU + cvf / N_CPU)
    rep["cases"].append({"product": cname, "l": l, "M": M, "scheme": scheme, "anti": anti,
                         "mean_dP_gpu": gd, "mean_dP_cpu": cd, "z_mean_dP": z_d,
                         "mean_Pf_gpu": gf, "mean_Pf_cpu": cf, "z_mean_Pf": z_f,
                         "var_dP_ratio": gvd / cvd, "var_Pf_ratio": gvf / cvf,
                         "seconds_cpu": t_cpu, "seconds_gpu": t_gpu})
    print(f"{cname:16s} l={l} M={M} {scheme:8s} anti={anti!s:5s} z(dP)={z_d:+.2f} z(Pf)={z_f:+.2f} "
          f"VdP ratio={gvd / cvd:.5f} VPf ratio={gvf / cvf:.5f}  cpu {t_cpu:.1f}s gpu {t_gpu:.2f}s", file=sys.stderr)
zs = [abs(c["z_mean_dP"]) for c in rep["cases"]] + [abs(c["z_mean_Pf"]) for c in rep["cases"]]
rep["max_abs_z"] = max(zs)
rep["n_z_above_3"] = int(sum(z > 3 for z in zs))
print(json.dumps(rep, indent=1))
