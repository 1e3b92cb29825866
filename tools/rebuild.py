#!/usr/bin/env python
"""Rebuild only some translation units of libmlmcb.so (kernel work): python tools/rebuild.py family5_s1.o family4_s1.o ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multilevel_mc_sdes_b200 import build as b  # noqa: E402

t0 = time.time()
only = [a for a in sys.argv[1:] if a != "-v"]
b.compile_lib(b.LIB, only=only or None, verbose="-v" in sys.argv)
b.write_sass_mix()
print(f"rebuilt {sys.argv[1:] or 'all'} in {time.time() - t0:.0f} s")
