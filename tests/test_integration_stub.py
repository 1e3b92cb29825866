"""CPU, build container only: the ctypes stub printed in INTEGRATION.md, executed against the real
reference module - it must map every reference product to the right (model, payoff) codes and reach the
C ABI (which, without a GPU, answers MLMCB_ECUDA: there is no CPU fallback)."""
import os
import re
import sys

import pytest

from multilevel_mc_sdes_b200 import _lib
from oracle.ref_loader import load_reference, reference_available
from tests.conftest import ROOT

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")


def test_integration_stub_runs_against_the_reference():
    ref = load_reference()
    sys.modules["MLMC_RSCAM"] = ref
    try:
        src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
        code = re.search(r"```python\n# mlmcb_plugin.py.*?\n(.*?)```", src, re.S).group(1)
        code = code.replace('C.CDLL("libmlmcb.so")', f'C.CDLL("{_lib.LIB_PATH}")')
        original = ref.Option.looper
        ns = {}
        exec(code, ns)
        import ctypes
        assert ctypes.sizeof(ns["_P"]) == ctypes.sizeof(_lib.Params)      # the stub's struct is the header's
        want = {"Euro_GBM": (0, 0), "Asian_GBM": (0, 1), "Lookback_GBM": (0, 2), "Digital_GBM": (0, 3),
                "Euro_Merton": (1, 0), "Asian_Merton": (1, 1), "Lookback_Merton": (1, 2), "Digital_Merton": (1, 3)}
        import torch
        for name, codes in want.items():
            opt = getattr(ref, name)()
            assert ns["_codes"](opt) == codes
            if not torch.cuda.is_available():
                with pytest.raises(RuntimeError, match="no CPU fallback"):
                    opt.looper(10, 1, 2, False)
    finally:
        ref.Option.looper = original
        sys.modules.pop("MLMC_RSCAM", None)
