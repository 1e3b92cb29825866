"""CPU, build container only: the plot-free Giles_plot core against the reference's own Giles_plot
run with recording axes, both consuming the same numpy stream through the reference's looper."""
import contextlib
import io

import numpy as np
import pytest

from multilevel_mc_sdes_b200 import giles_analysis
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")


class Axis:
    def __init__(self):
        self.calls = []
        self.xaxis = self

    def _rec(self, kind, *a):
        self.calls.append((kind, [np.array(x, dtype=float, copy=True) for x in a[:2]]))

    def plot(self, *a, **k):
        self._rec("plot", *a)

    def semilogy(self, *a, **k):
        self._rec("semilogy", *a)

    def loglog(self, *a, **k):
        self._rec("loglog", *a)

    def get_legend_handles_labels(self):
        return [], []

    def __getattr__(self, name):          # set_xlabel, legend, annotate, add_artist, set_major_locator ...
        return lambda *a, **k: None


class Fig:
    def __init__(self):
        self.axes = [Axis() for _ in range(4)]

    def get_size_inches(self):
        return [10.0, 10.0]

    def suptitle(self, *a, **k):
        pass

    def tight_layout(self, *a, **k):
        pass


@pytest.mark.parametrize("cname,kw,anti,M", [("Euro_GBM", dict(alpha_0=1, beta_0=1), True, 2),
                                             ("Lookback_GBM", dict(alpha_0=1, beta_0=1), False, 4),
                                             ("Digital_Merton", dict(alpha_0=1, beta_0=0.5), False, 2)])
def test_giles_core_matches_reference_plot_data(cname, kw, anti, M):
    ref = load_reference()
    eps = [0.5, 0.3]
    Lmax, Nsamples = (4, 2000) if "Merton" not in cname else (3, 300)
    fig = Fig()
    np.random.seed(42)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        ref.Giles_plot(getattr(ref, cname)(**kw), eps, ["o", "s"], "x", fig, M=M, N0=200, anti=anti, Lmax=Lmax,
                       Nsamples=Nsamples)
    np.random.seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        out = giles_analysis(getattr(ref, cname)(**kw), eps, M=M, N0=200, anti=anti, Lmax=Lmax, Nsamples=Nsamples)
    var_ax, mean_ax, n_ax, cost_ax = fig.axes
    eq = lambda a, b: np.array_equal(a, b, equal_nan=True)
    assert eq(var_ax.calls[0][1][1], out["log_var_P"])
    assert eq(var_ax.calls[1][1][1], out["log_var_dP"])
    assert eq(mean_ax.calls[0][1][1], out["log_mean_P"])
    assert eq(mean_ax.calls[1][1][1], out["log_mean_dP"])
    if anti:
        assert eq(var_ax.calls[2][1][1], out["log_var_dP_anti"])
        assert eq(mean_ax.calls[2][1][1], out["log_mean_dP_anti"])
    n_calls = [c[1][1] for c in n_ax.calls]
    if anti:
        assert all(eq(a, b) for a, b in zip(n_calls[0::2], out["N"]))
        assert all(eq(a, b) for a, b in zip(n_calls[1::2], out["N_anti"]))
    else:
        assert all(eq(a, b) for a, b in zip(n_calls, out["N"]))
    assert eq(cost_ax.calls[0][1][1], out["cost_mc"])
    assert eq(cost_ax.calls[1][1][1], out["cost_mlmc"])
    if anti:
        assert eq(cost_ax.calls[2][1][1], out["cost_anti"])
