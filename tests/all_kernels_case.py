#!/usr/bin/env python
"""Small end-to-end script (run by hand on a GPU box, also the intended target of a memcheck run):
every kernel family once at small, non-multiple-of-block sizes.  Lives under tests/ because it uses the
oracle's draw recorder."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402
from oracle import np_oracle as O  # noqa: E402

for cls in (mm.Euro_GBM, mm.Asian_GBM, mm.Lookback_GBM, mm.Digital_GBM, mm.Euro_Merton, mm.Asian_Merton,
            mm.Lookback_Merton, mm.Digital_Merton):
    # GBM: packed kernel (1, 2 steps) and level kernel; Merton: l = 0 loop, event-driven (2, 16, 9 steps), block-uniform (256)
    for (l, M) in ((0, 2), (1, 2), (2, 4), (2, 3), (8, 2)):
        o = cls(seed=1)
        s = o.looper(3001, l, M, False)
        assert np.isfinite(s).all()
    Pf, Pc = cls(seed=2).payoff(777, 2, 4)
    assert np.isfinite(Pf).all()
    assert np.isfinite(cls(seed=3, scheme="milstein").looper(2000, 2, 4, False)).all()
    assert np.isfinite(cls(seed=3, sampler="f64bm").looper(2000, 2, 2, False)).all()
    assert np.isfinite(cls(seed=3, fp32=True).looper(2000, 2, 2, False)).all()
for cls in (mm.Euro_GBM, mm.Asian_GBM):
    assert np.isfinite(cls(seed=4).looper(2500, 3, 4, True)).all()
    assert np.isfinite(cls(seed=4).looper(2500, 1, 6, True)).all()
Z = np.random.RandomState(0).randn(16, 100)
for cls in (mm.Euro_GBM, mm.Asian_GBM, mm.Lookback_GBM):
    for ar in ("exact", "fast"):
        cls().payoff_replay(Z, 2, 4, arith=ar)
mm.Euro_GBM().payoff_replay(Z, 2, 4, anti=True)
rec = O.RecordingDraws()
np.random.seed(1)
O.level_payoffs(O.Params(), O.MERTON, O.ASIAN, 2, 4, n=50, draws=rec)
mm.Asian_Merton().payoff_replay(rec.packed(), 2, 4, arith="exact")
sw = mm.Lookback_GBM().level_sweep(Lmax=4, Nsamples=5000, M=2)
assert sw.shape == (7, 5)
print("sanitize case ok")
