import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CUDA library and the C oracle are build products (git-ignored): make sure they exist and
    are current before any test imports them (nvcc cross-compiles sm_100a without a GPU)."""
    import warnings
    from multilevel_mc_sdes_b200 import build as b
    if not os.path.exists(b.LIB):
        try:
            b.build()
        except Exception as e:  # noqa: BLE001 - no nvcc here: the tests that need the library fail on their own, the
            warnings.warn(f"libmlmcb.so is missing and could not be built: {e}")      # oracle / driver-logic tests still run
    elif b.is_stale():
        try:            # file times may not survive a copy of the tree: a present library is still usable
            b.build()
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"libmlmcb.so looks older than its sources and could not be rebuilt: {e}")
    from oracle import c_oracle
    try:
        c_oracle.build()
    except Exception:  # noqa: BLE001
        if not os.path.exists(os.path.join(ROOT, "oracle", "libmlmc_oracle.so")):
            raise


@pytest.fixture(scope="session")
def golden():
    return {n: np.load(os.path.join(GOLDEN, n + ".npz"), allow_pickle=False)
            for n in ("payoffs_gbm", "payoffs_merton", "payoffs_merton_deep", "payoffs_anti", "payoffs_anti_milstein", "payoffs_milstein", "payoffs_amer", "payoffs_amer_callable", "looper", "mlmc_trace", "closed_forms")}


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a real device; fails (not skips) if it is not there."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from multilevel_mc_sdes_b200 import _lib
    return _lib.load()
