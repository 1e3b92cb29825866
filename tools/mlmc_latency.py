#!/usr/bin/env python
"""Where the wall time of a latency-bound mlmc run goes: every sweep of Euro_GBM(125,105,.02,.5).mlmc(0.01) with its
requests and its wall time, the host-side share (driver arithmetic between sweeps), best of a few runs.

    python tools/mlmc_latency.py            # fused sweeps (one launch per iteration)
    MLMCB_NO_FUSED_SWEEP=1 python tools/mlmc_latency.py     # per-level launches on side streams
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_mc_sdes_b200 as mm  # noqa: E402

best = None
for rep in range(5):
    o = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, verbose=False, seed=rep)
    log = []
    inner = o._sweep

    def sweep(reqs, M, anti=False, _inner=inner, _log=log):
        t0 = time.perf_counter()
        out = _inner(reqs, M, anti=anti)
        _log.append((time.perf_counter() - t0, list(reqs)))
        return out

    o._sweep = sweep
    t0 = time.perf_counter()
    sums, N = o.mlmc(0.01)
    wall = time.perf_counter() - t0
    if best is None or wall < best[0]:
        best = (wall, log, len(N))
wall, log, levels = best
dev = sum(t for t, _ in log)
print(f"wall {1e3 * wall:.2f} ms, {len(log)} sweeps, {levels} levels; inside sweeps {1e3 * dev:.2f} ms, host driver {1e3 * (wall - dev):.2f} ms")
for t, reqs in log:
    deepest = max(l for _, l in reqs)
    print(f"  {1e3 * t:7.3f} ms  levels {len(reqs):2d}  deepest l={deepest:2d}  samples {sum(n for n, _ in reqs):.3g}  at l=0: {dict((l, n) for n, l in reqs).get(0, 0):.3g}")
