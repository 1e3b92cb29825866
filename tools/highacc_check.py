#!/usr/bin/env python
"""High-accuracy end-to-end check of both samplers: Euro_GBM and Euro_Merton mlmc at small eps against the
reference's closed forms BS() (Black-Scholes :1383-1393, Merton's series :1648-1664)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402

rep = {}
for sampler in ("f32bm", "f64bm"):
    for eps in ((2e-3, 5e-4, 1e-4, 5e-5) if sampler == "f32bm" else (5e-4, 1e-4)):
        errs, secs, tot = [], [], []
        for seed in range(4):
            o = mm.Euro_GBM(verbose=False, seed=seed, sampler=sampler, alpha_0=1, beta_0=1)
            t0 = time.perf_counter()
            sums, N = o.mlmc(eps, M=4)
            secs.append(time.perf_counter() - t0)
            errs.append(float(np.sum(sums[0] / N) - o.BS()))
            tot.append(float(N.sum()))
        rep[f"{sampler}_eps{eps:g}"] = {"errors": errs, "rms_over_eps": float(np.sqrt(np.mean(np.square(errs))) / eps),
                                        "mean_err_over_eps": float(np.mean(errs) / eps), "seconds": secs, "samples": tot}
# Merton: every walk takes part (l = 0 lane-queue loop, event-driven below 128 steps, block-uniform above)
for sampler, M, epss in (("f32bm", 2, (2e-3, 5e-4)), ("f32bm", 4, (2e-3, 5e-4)), ("f64bm", 2, (2e-3,))):
    for eps in epss:
        errs, secs, tot, Ls = [], [], [], []
        for seed in range(4):
            o = mm.Euro_Merton(verbose=False, seed=100 + seed, sampler=sampler, alpha_0=1, beta_0=1)
            t0 = time.perf_counter()
            sums, N = o.mlmc(eps, M=M)
            secs.append(time.perf_counter() - t0)
            errs.append(float(np.sum(sums[0] / N) - o.BS()))
            tot.append(float(N.sum()))
            Ls.append(len(N) - 1)
        rep[f"merton_{sampler}_M{M}_eps{eps:g}"] = {"errors": errs, "rms_over_eps": float(np.sqrt(np.mean(np.square(errs))) / eps),
                                                    "mean_err_over_eps": float(np.mean(errs) / eps), "seconds": secs,
                                                    "samples": tot, "L": Ls}
print(json.dumps(rep, indent=1))
