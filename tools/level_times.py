#!/usr/bin/env python
"""Per-level cost table of one product: ms per launch and SMSP cycles per warp-path for l = 0..Lmax.

    python tools/level_times.py --cls Digital_Merton --M 2 --Lmax 8 [--flags 0]
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
ap.add_argument("--cls", default="Digital_Merton")
ap.add_argument("--M", type=int, default=2)
ap.add_argument("--Lmax", type=int, default=8)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--n", type=float, default=0.0, help="samples per launch; default: 2^32 / steps clamped to [2^24, 2^30]")
a = ap.parse_args()
lib = _lib.load()
opt = getattr(mm, a.cls)()
p = opt._params()
out = torch.zeros(7, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for l in range(a.Lmax + 1):
    n = int(a.n) if a.n else max(1 << 24, min(1 << 30, (1 << 32) // a.M ** l))
    best = 1e9
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mlmcb_level_sums(C.byref(p), opt._model, opt._payoff_kind, l, a.M, n, 0, r * n, a.flags, 0, st,
                                        out.data_ptr(), None, None))
        e1.record()
        torch.cuda.synchronize()
        if r:
            best = min(best, e0.elapsed_time(e1))
    cyc = 148 * 4 * 1.965e9 * best * 1e-3 / (n / 32)
    print(f"{a.cls} M={a.M} l={l} steps={a.M ** l:5d}: {best:8.3f} ms / {n:.2e} samples  {n / best * 1e3:.3e} samples/s  "
          f"{cyc:8.1f} cyc/warp-path  {cyc / a.M ** l:7.1f} cyc/warp-step  mean dP={out[0].item() / n:+.6f}")
