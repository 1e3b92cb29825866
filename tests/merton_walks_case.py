"""Helper for test_gpu_philox.py::test_merton_walks_agree_per_path (run as a subprocess, because the walk
selection is read from the environment once per process): per-path payoffs of the Merton products at a few
levels, written to the .npz given as argv[1]."""
import sys

import numpy as np

import multilevel_mc_sdes_b200 as mm

out = {}
for cname in ("Euro_Merton", "Asian_Merton", "Lookback_Merton", "Digital_Merton"):
    for l, M, kw in ((1, 2, {}), (2, 2, {}), (3, 2, {}), (2, 4, {}), (2, 3, {}), (6, 2, {"lam": 5.0}), (3, 4, {"scheme": "milstein"}),
                     (3, 2, {"lam": 40.0}), (8, 2, {"lam": 12.0}),         # many more jumps than records kept ahead
                     (2, 2, {"lam": 300.0, "jumpstd": 0.02, "jumpmean": -0.0002})):   # dozens of record rounds inside one fine step
        for sampler in ("f64bm", "f32bm"):
            opt = getattr(mm, cname)(seed=23, sampler=sampler, **kw)
            opt._next_path[l] = 77
            Pf, Pc = opt.payoff(4099 if kw.get("lam", 1.0) < 100 else 515, l, M)
            key = f"{cname}_{l}_{M}_{sorted(kw.items())}_{sampler}"
            out[key + "_f"] = Pf
            out[key + "_c"] = Pc
np.savez(sys.argv[1], **out)
