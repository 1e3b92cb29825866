#!/usr/bin/env python
"""One-off fuzz of the Merton walks (run by hand on a GPU box): N random parameter sets / levels / path-id
offsets through the own-draws replay comparison of tests/test_gpu_philox.py.  Lives under tests/ because the
comparison goes through the oracle-checked replay path.

    python tests/merton_fuzz.py [N=150] [seed=1]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402
from multilevel_mc_sdes_b200 import _lib  # noqa: E402
from tests.test_gpu_philox import _merton_own_draws_case  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
lib = _lib.load()
classes = (mm.Euro_Merton, mm.Asian_Merton, mm.Lookback_Merton, mm.Digital_Merton)
levels = ((0, 2), (0, 4), (1, 2), (2, 2), (1, 4), (3, 2), (2, 3), (1, 5), (2, 4), (5, 2), (3, 4), (7, 2), (4, 4), (9, 2))
done = 0
for it in range(N):
    kw = dict(X0=float(rng.uniform(1, 500)), T=float(rng.uniform(0.05, 5.0)), r=float(rng.uniform(-0.02, 0.15)),
              sig=float(rng.uniform(0.01, 0.9)), lam=float(rng.choice([0.05, 0.5, 1.0, 3.0, 10.0, 30.0])),
              jumpmean=float(rng.uniform(-0.5, 0.5)), jumpstd=float(rng.uniform(0.01, 0.8)))
    kw["K"] = kw["X0"] * float(rng.uniform(0.5, 1.5))
    cls = classes[int(rng.integers(0, 4))]
    l, M = levels[int(rng.integers(0, len(levels)))]
    lamT = kw["lam"] * kw["T"]
    K = int(min(4096, max(64, 6 * lamT + 60)))
    n = 200 if M ** l <= 128 else 60
    try:
        _merton_own_draws_case(lib, cls, kw, l, M, n=n, off=int(rng.integers(0, 2 ** 44)), seed=int(rng.integers(0, 2 ** 31)), K=K)
    except AssertionError as e:
        print("FAIL", it, cls.__name__, l, M, kw, str(e)[:300])
        sys.exit(1)
    done += 1
print(f"merton fuzz: {done} cases ok")
