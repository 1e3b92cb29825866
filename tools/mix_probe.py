#!/usr/bin/env python
"""The headline loop's instruction mix issued by one kind of warp or by specialised warps (mlmcb_mix_probe):
SMSP cycles per fine step of one warp."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multilevel_mc_sdes_b200 import _lib  # noqa: E402

lib = _lib.load()
iters = 200000
for mode, name in ((0, "every warp the whole mix (2 warps per SMSP)"), (1, "2 FP64 warps + 1 Philox warp per SMSP")):
    ms = C.c_float()
    _lib.check(lib.mlmcb_mix_probe(mode, 0, iters, C.byref(ms)))
    # per SMSP one pass is two warp-steps of the mix in both modes
    print(f"mode {mode}: {name}: {ms.value:.2f} ms -> {ms.value * 1e-3 * 1.965e9 / iters / 2:.1f} SMSP cycles per warp-step")
