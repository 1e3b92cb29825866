"""Load the upstream reference (`/root/reference/MLMC_RSCAM.py`) in THIS container only.

TEST INFRASTRUCTURE. Not shipped, not imported by the product package. Used by
`oracle/make_golden.py` (fixture generation) and by the CPU tests that are skipped
when `/root/reference` is absent (the GPU box has no reference tree).

The reference does `import matplotlib.pyplot` at MLMC_RSCAM.py:34-38 and matplotlib is
not installed here, so inert stand-in modules are put in `sys.modules` before the
import. Nothing on the sampler path touches them.
"""
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_PY = os.environ.get("MLMC_REFERENCE_PY", "/root/reference/MLMC_RSCAM.py")


def reference_available():
    return os.path.isfile(REFERENCE_PY)


def _stub_matplotlib():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_mlmcb_stub", False):
        return  # a real matplotlib is importable; leave it alone
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    root = types.ModuleType("matplotlib")
    root._mlmcb_stub = True
    for sub in ("pyplot", "ticker", "legend"):
        m = types.ModuleType("matplotlib." + sub)
        setattr(root, sub, m)
        sys.modules["matplotlib." + sub] = m
    # inert stand-ins for the few names Giles_plot touches (:1829, :1862, :1870-1873), so the
    # reference's own function can be run against recording axes in tests/test_giles_core.py

    class Legend:
        def __init__(self, *a, **k):
            self._legend_box = types.SimpleNamespace()

    sys.modules["matplotlib.legend"].Legend = Legend
    sys.modules["matplotlib.ticker"].MaxNLocator = lambda *a, **k: None
    sys.modules["matplotlib.pyplot"].Line2D = lambda *a, **k: None
    sys.modules["matplotlib"] = root


_cached = None


def load_reference():
    """Return the reference module object (imported once, unmodified)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(REFERENCE_PY)
    _stub_matplotlib()
    spec = importlib.util.spec_from_file_location("MLMC_RSCAM_reference", REFERENCE_PY)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarnings at :1870-1886 (plot labels)
        spec.loader.exec_module(mod)
    _cached = mod
    return mod
