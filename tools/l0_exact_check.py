#!/usr/bin/env python
"""Sharp end-to-end check of the normal samplers through the product path: at l = 0 a GBM path is ONE Euler
step, X = X0(1 + rT) + X0 sig sqrt(T) z, so E[e^{-rT} max(X-K, 0)] is known in closed form (normal X) and
the level-0 mean of `looper` over ~10^12 samples resolves a bias of 10^-6 in the mean of z (or 3*10^-6 in
its variance) - far below what the moment checks of tools/sampler_quality.py can see.

    python tools/l0_exact_check.py [--samples 4e11] [--seeds 8] > gpurun_out/l0_exact.json
"""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=float, default=4e11)
ap.add_argument("--seeds", type=int, default=8)
a = ap.parse_args()
Phi = lambda x: 0.5 * math.erfc(-x / math.sqrt(2.0))
phi = lambda x: math.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
rep = {}
for name, kw in (("Euro_GBM", {}), ("Euro_GBM_itm", dict(X0=125, K=105, r=0.02, sig=0.5)), ("Digital_GBM", {})):
    cls = getattr(mm, name.split("_itm")[0])
    o = cls(verbose=False, **kw)
    m, s = o.X0 * (1 + o.r * o.T), o.X0 * o.sig * math.sqrt(o.T)
    d = (m - o.K) / s
    disc = math.exp(-o.r * o.T)
    if name.startswith("Digital"):
        exact, second = disc * o.K * Phi(d), (disc * o.K) ** 2 * Phi(d)
    else:
        exact = disc * ((m - o.K) * Phi(d) + s * phi(d))
        second = disc ** 2 * (((m - o.K) ** 2 + s * s) * Phi(d) + (m - o.K) * s * phi(d))
    var = second - exact ** 2
    for sampler, n in (("f64bm", int(a.samples)), ("f32bm", int(a.samples))):
        means = []
        for seed in range(a.seeds):
            sums = cls(verbose=False, seed=1000 + seed, sampler=sampler, **kw).looper(n, 0, 2, False)
            means.append(sums[2] / n)
        means = np.array(means)
        se1 = math.sqrt(var / n)
        rep[f"{name}_{sampler}"] = {"exact": exact, "samples_per_seed": n, "means": means.tolist(),
                                    "z_per_seed": ((means - exact) / se1).tolist(),
                                    "mean_error": float(means.mean() - exact), "se_of_mean_error": se1 / math.sqrt(a.seeds),
                                    "z": float((means.mean() - exact) / (se1 / math.sqrt(a.seeds)))}
print(json.dumps(rep, indent=1))
