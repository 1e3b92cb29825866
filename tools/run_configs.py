#!/usr/bin/env python
"""Run the BASELINE.json configurations on one B200 and print a JSON report.

    python tools/run_configs.py [--quick] > gpurun_out/configs.json

#1 Euro_GBM(125,105,.02,.5).mlmc(0.01) vs BS()          (README example)
#2 Asian_GBM().mlmc(1e-4, M=4), fp64, alpha_0=beta_0=None
#3 Lookback_GBM() fixed-level sweep l=0..8, 1e8 samples/level, M=2 and M=4 (Giles_plot core :1775-1777)
#4 Digital_Merton(beta_0=0.5).mlmc(1e-3, M=2)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import multilevel_mc_sdes_b200 as mm  # noqa: E402


def timed_mlmc(opt, eps, **kw):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sums, N = opt.mlmc(eps, **kw)
    dt = time.perf_counter() - t0
    L = len(N) - 1
    M = kw.get("M", 2)
    return {"eps": eps, "seconds": dt, "price": float(np.sum(sums[0] / N)), "L": L, "N": [float(n) for n in N],
            "total_samples": float(N.sum()), "fine_steps": float(np.sum(N * M ** np.arange(L + 1))),
            "alpha": opt.alpha_0, "beta": opt.beta_0,
            "var_dP": [float(v) for v in (sums[1] / N - (sums[0] / N) ** 2)]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    rep = {"gpu": torch.cuda.get_device_name(0)}
    mm.Euro_GBM(verbose=False).looper(1000, 1, 2, False)          # context + module load outside the timings

    o = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, verbose=False)
    r = timed_mlmc(o, 0.01)
    r["BS"] = float(o.BS())
    r["abs_err"] = abs(r["price"] - r["BS"])
    rep["cfg1_readme_euro_gbm_eps0.01"] = r

    o = mm.Asian_GBM(verbose=False)
    eps = 1e-3 if a.quick else 1e-4
    rep[f"cfg2_asian_gbm_M4_eps{eps:g}"] = timed_mlmc(o, eps, M=4)

    nlev = 10 ** 7 if a.quick else 10 ** 8
    for M in (2, 4):
        o = mm.Lookback_GBM(verbose=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s = o.level_sweep(Lmax=8, Nsamples=nlev, M=M)
        dt = time.perf_counter() - t0
        mean_dp = np.abs(s[0] / nlev)
        var_dp = s[1] / nlev - mean_dp ** 2
        X = np.ones((8, 2))
        X[:, 0] = np.arange(1, 9)
        alpha = -np.linalg.lstsq(X, np.log(mean_dp[1:]), rcond=None)[0][0] / np.log(M)      # :1820-1823
        beta = -np.linalg.lstsq(X, np.log(var_dp[1:]), rcond=None)[0][0] / np.log(M)
        steps = float(nlev * np.sum(M ** np.arange(9)))
        rep[f"cfg3_lookback_sweep_M{M}"] = {"samples_per_level": nlev, "seconds": dt, "fine_steps": steps,
                                           "steps_per_s": steps / dt, "alpha": float(alpha), "beta": float(beta),
                                           "log_M_var_dP": [float(v) for v in np.log(var_dp[1:]) / np.log(M)],
                                           "price_telescoped": float(np.sum(s[0] / nlev)), "BS": float(o.BS())}

    o = mm.Digital_Merton(beta_0=0.5, verbose=False)
    eps = 1e-2 if a.quick else 1e-3
    rep[f"cfg4_digital_merton_M2_eps{eps:g}"] = timed_mlmc(o, eps, M=2)

    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
