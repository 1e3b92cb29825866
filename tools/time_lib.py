#!/usr/bin/env python
"""Time one level call of a given build of the library (kernel experiments).

    MLMCB_LIB=/path/lib.so python tools/time_lib.py [--payoff 0] [--M 4] [--l 6] [--model 0] [--sampler 0]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multilevel_mc_sdes_b200 import _lib  # noqa: E402
import multilevel_mc_sdes_b200 as mm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", type=int, default=0)
ap.add_argument("--payoff", type=int, default=0)
ap.add_argument("--M", type=int, default=4)
ap.add_argument("--l", type=int, default=6)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--batch", type=float, default=float(1 << 25))
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
lib = _lib.load()
cls = {(0, 4): mm.Amer_GBM, (0, 0): mm.Euro_GBM, (0, 1): mm.Asian_GBM, (0, 2): mm.Lookback_GBM, (0, 3): mm.Digital_GBM,
       (1, 0): mm.Euro_Merton, (1, 1): mm.Asian_Merton, (1, 2): mm.Lookback_Merton, (1, 3): mm.Digital_Merton}[(a.model, a.payoff)]
opt = cls(X0=125, K=105, r=0.02, sig=0.5) if (a.model == 0 and a.payoff != 4) else cls()
p = opt._params()
out = torch.zeros(7, dtype=torch.float64, device="cuda")
n = int(a.batch)
st = torch.cuda.current_stream().cuda_stream


def go(off):
    _lib.check(lib.mlmcb_level_sums(C.byref(p), a.model, a.payoff, a.l, a.M, n, 0, off, a.flags, 0, st, out.data_ptr(), None, None))


go(0)
torch.cuda.synchronize()
best = 1e9
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go((r + 1) * n)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
s = out.cpu().numpy()
steps = n * a.M ** a.l
print(f"{os.environ.get('MLMCB_LIB', 'default'):40s} model={a.model} payoff={a.payoff} l={a.l} M={a.M} flags={a.flags}: "
      f"{best:8.2f} ms  {n / best * 1e3:.4e} samples/s  {steps / best * 1e3:.4e} steps/s  "
      f"cyc/warp-step/SMSP={148 * 4 * 1.965e9 / (steps / best * 1e3 / 32):.1f}  meanPf={s[2] / n:.5f}")
