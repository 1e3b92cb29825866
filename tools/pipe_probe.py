#!/usr/bin/env python
"""Print the per-pipe issue rates measured by mlmcb_pipe_probe (roofline denominators)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multilevel_mc_sdes_b200 import _lib  # noqa: E402

NAMES = ["DFMA", "DADD", "FFMA", "IMAD.WIDE", "LOP3", "MUFU.LG2", "I2F.U32", "F2F(f32<->f64)", "MUFU.SIN",
         "DFMA+LOP3", "DFMA+IMAD.WIDE", "IMAD.WIDE+LOP3", "FFMA+LOP3", "DFMA+FFMA", "DFMA+IMAD", "IMAD.HI", "DFMA+IMAD.HI",
         "IMAD.HI+LOP3", "IMAD.HI+IMAD"]


def main():
    lib = _lib.load()
    res = {}
    for which, name in enumerate(NAMES):
        best = 0.0
        for _ in range(3):
            ips, ms = C.c_double(), C.c_float()
            _lib.check(lib.mlmcb_pipe_probe(which, 0, C.byref(ips), C.byref(ms)))
            best = max(best, ips.value)
        res[name] = {"warp_inst_per_s": best, "lane_ops_per_s": best * 32,
                     "per_sm_per_clk_at_1965MHz": best * 32 / 148 / 1.965e9}
        n_inst = 2 if "+" in name or name.startswith("F2F") else 1
        print(f"{name:16s} {best:.4e} warp-inst/s   {best * 32 / 148 / 1.965e9:7.2f} lanes/SM/clk @1965MHz   "
              f"{148 * 4 * 1.965e9 / best * n_inst:5.2f} SMSP cycles per {'pair' if n_inst == 2 else 'instruction'}")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
