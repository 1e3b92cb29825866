"""GPU: end-to-end Option.mlmc(eps) against the reference's closed forms (north-star check (b))."""
import numpy as np
import pytest

import multilevel_mc_sdes_b200 as mm

pytestmark = pytest.mark.gpu


def price(sums, N):
    return float(np.sum(sums[0, :] / N))          # README.md:29


def test_readme_example_within_eps_of_BS(cuda_lib):
    """BASELINE config #1: Euro_GBM(125,105,.02,.5).mlmc(0.01) vs BS() = 35.17419437739385, with
    alpha_0 = beta_0 = None exactly as the README writes it."""
    errs = []
    for seed in range(5):
        opt = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5, verbose=False, seed=seed)
        sums, N = opt.mlmc(eps=0.01)
        assert sums.shape == (7, len(N)) and N.dtype == np.float64
        errs.append(abs(price(sums, N) - opt.BS()))
        assert opt.alpha_0 is not None and opt.beta_0 is not None  # warm_start side effect (:1037)
    # eps is an RMS target (half bias^2, half variance, :1021/:1025): the typical run is inside
    # eps, no run is far outside
    assert np.median(errs) < 0.01 and max(errs) < 0.03, errs


def test_rms_error_over_seeds_matches_eps(cuda_lib):
    """mlmc targets an RMS error of eps (eps^2/2 variance + eps^2/2 squared bias, :1021/:1025):
    over 24 independent seeds the RMS error is ~eps (sampling slack 35 %) and the mean error,
    i.e. the bias, stays inside eps/sqrt(2) plus its own standard error."""
    eps, errs = 0.02, []
    for seed in range(24):
        opt = mm.Euro_GBM(seed=100 + seed, verbose=False)
        sums, N = opt.mlmc(eps=eps, M=4)
        errs.append(price(sums, N) - opt.BS())
    errs = np.array(errs)
    assert np.sqrt(np.mean(errs ** 2)) < 1.35 * eps, errs
    assert abs(errs.mean()) < eps / np.sqrt(2) + 3 * errs.std() / np.sqrt(len(errs)), errs


@pytest.mark.parametrize("cls,kw,eps,M", [
    (mm.Lookback_GBM, {}, 0.02, 4),
    (mm.Digital_GBM, dict(beta_0=0.5), 0.05, 4),
    (mm.Euro_Merton, {}, 0.03, 2),
    (mm.Euro_GBM, dict(sampler="f64bm"), 0.02, 2),
])
def test_other_closed_forms(cuda_lib, cls, kw, eps, M):
    opt = cls(verbose=False, **kw)
    sums, N = opt.mlmc(eps=eps, M=M)
    assert abs(price(sums, N) - opt.BS()) < 2 * eps


def test_asian_matches_reference_limit(cuda_lib):
    # Asian BS() is the Turnbull-Wakeman approximation (5.7828); MLMC converges to ~5.763
    opt = mm.Asian_GBM(verbose=False)
    sums, N = opt.mlmc(eps=0.005, M=4)
    assert abs(price(sums, N) - 5.7626) < 0.012


def test_mlmc_antithetic(cuda_lib):
    for cls in (mm.Euro_GBM, mm.Asian_GBM):
        opt = cls(verbose=False, seed=3)
        sums, N = opt.mlmc(eps=0.01, M=4, anti=True)
        ref = opt.BS() if cls is mm.Euro_GBM else 5.7626
        assert abs(price(sums, N) - ref) < 0.03
    with pytest.raises(ValueError, match="odd coarseness factor"):
        mm.Euro_GBM(verbose=False).mlmc(eps=0.1, M=3, anti=True)


def test_giles_analysis_on_gpu(cuda_lib):
    """BASELINE config #3 shape: fixed-level sweep + regressions (Giles_plot core :1741-1823)."""
    opt = mm.Lookback_GBM(verbose=False, alpha_0=1, beta_0=1)
    out = mm.giles_analysis(opt, [0.1, 0.05], M=4, Lmax=6, Nsamples=2_000_000)
    assert out["log_var_dP"].shape == (6,) and len(out["N"]) == 2
    assert 0.7 < out["alpha"] < 1.3 and 0.7 < out["beta"] < 1.3       # Euler-Maruyama, lookback with offset
    assert np.all(np.diff(out["log_var_dP"]) < 0)
    assert out["cost_mlmc"][1] < out["cost_mc"][1]                    # MLMC beats plain MC at the finer eps
    euro = mm.Euro_GBM(verbose=False, alpha_0=1, beta_0=1)
    out = mm.giles_analysis(euro, [0.1], M=2, anti=True, Lmax=5, Nsamples=1_000_000)
    assert out["log_var_dP_anti"].shape == (5,) and len(out["N_anti"]) == 1


def _stand_in_reference():
    """A minimal MLMC_RSCAM look-alike for the GPU box (the reference itself is not there): the free path / payoff
    functions the reference binds as class attributes (:1378-1381, :1405-1408, :1442-1444, :1470-1471, :1645-1704),
    the class tree, and instances carrying the reference's attribute names and default values."""
    import types
    ref = types.ModuleType("MLMC_RSCAM")

    def host_only(*a, **k):
        raise AssertionError("the reference's host path must not run once looper is replaced")
    for name in ("EA_payoff", "Lookback_payoff", "Digital_payoff", "diffusion_path", "diffusion_path_avg",
                 "diffusion_path_min", "JD_path", "JD_path_avg", "JD_path_min"):
        setattr(ref, name, types.FunctionType(host_only.__code__, {}, name))

    class Option:
        def __init__(self, alpha_0=None, beta_0=None, X0=100, K=100, T=1, r=0.05):
            self.alpha_0, self.beta_0, self.X0, self.K, self.T, self.r = alpha_0, beta_0, X0, K, T, r

        def looper(self, Nl, l, M, anti, Npl=10 ** 4):
            raise AssertionError("Option.looper was not replaced")

    class GBM_Option(Option):
        def __init__(self, drift=0, sig=0.2, **kw):
            super().__init__(**kw)
            self.sig, self.mu = sig, self.r + drift

    class Merton_Option(Option):
        def __init__(self, lam=1, jumpmean=0.1, jumpstd=0.2, sig=0.2, **kw):
            super().__init__(**kw)
            self.lam, self.jumpmean, self.jumpstd, self.sig = lam, jumpmean, jumpstd, sig
            self.J_bar = np.exp(jumpmean + 0.5 * jumpstd ** 2) - 1

    ref.Option, ref.GBM_Option, ref.Merton_Option = Option, GBM_Option, Merton_Option
    for name, base, path, payoff, extra in (
            ("Euro_GBM", GBM_Option, "diffusion_path", "EA_payoff", {}), ("Asian_GBM", GBM_Option, "diffusion_path_avg", "EA_payoff", {}),
            ("Lookback_GBM", GBM_Option, "diffusion_path_min", "Lookback_payoff", {"beta": 0.5826}),
            ("Digital_GBM", GBM_Option, "diffusion_path", "Digital_payoff", {}),
            ("Euro_Merton", Merton_Option, "JD_path", "EA_payoff", {}), ("Asian_Merton", Merton_Option, "JD_path_avg", "EA_payoff", {}),
            ("Lookback_Merton", Merton_Option, "JD_path_min", "Lookback_payoff", {"beta": 0.5826}),
            ("Digital_Merton", Merton_Option, "JD_path", "Digital_payoff", {})):
        setattr(ref, name, type(name, (base,), dict(path=getattr(ref, path), payoff=getattr(ref, payoff), **extra)))
    return ref


def test_integration_stub_on_the_gpu(cuda_lib):
    """The ctypes plug-in printed in INTEGRATION.md, executed on a B200: `Option.looper = gpu_looper` on a stand-in
    for the reference module gives, for every product, the very sums of this package's class with the same seed
    (same C-ABI entry point family, same path ids), consumes fresh ids on every call, and carries an `mlmc` run."""
    import os
    import re
    import sys
    from multilevel_mc_sdes_b200 import _lib
    from multilevel_mc_sdes_b200.options import mlmc_driver
    from tests.conftest import ROOT
    ref = _stand_in_reference()
    sys.modules["MLMC_RSCAM"] = ref
    try:
        src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
        code = re.search(r"```python\n# mlmcb_plugin.py.*?\n(.*?)```", src, re.S).group(1)
        code = code.replace('C.CDLL("libmlmcb.so")', f'C.CDLL("{_lib.LIB_PATH}")')
        ns = {}
        exec(code, ns)
        assert ref.Option.looper is ns["gpu_looper"]
        for name in ("Euro_GBM", "Asian_GBM", "Lookback_GBM", "Digital_GBM", "Euro_Merton", "Asian_Merton",
                     "Lookback_Merton", "Digital_Merton"):
            for (l, M, n) in ((0, 2, 100_003), (3, 4, 50_000), (7, 2, 20_000)):
                theirs, ours = getattr(ref, name)(), getattr(mm, name)(seed=0, verbose=False)
                a, b = theirs.looper(n, l, M, False), ours.looper(n, l, M, False)
                assert a.shape == (7,) and np.array_equal(a, b), (name, l, M, a, b)
                a2 = theirs.looper(n, l, M, False)                      # second call: fresh path ids, like ours
                assert not np.array_equal(a2, a) and np.array_equal(a2, ours.looper(n, l, M, False))
        anti = ref.Euro_GBM().looper(40_000, 2, 4, True)
        assert np.array_equal(anti, mm.Euro_GBM(seed=0).looper(40_000, 2, 4, True))
        # the reference's own driver loop shape, fed by the plug-in (the shipped `mlmc` needs alpha_0/beta_0 given, :982)
        opt = ref.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5)
        sums, N, _, _ = mlmc_driver(lambda reqs, M_: [opt.looper(n, l, M_, False) for n, l in reqs], 0.02, 2, 1000, 1, 1)
        assert abs(float(np.sum(sums[0] / N)) - 35.17419437739385) < 0.06
    finally:
        sys.modules.pop("MLMC_RSCAM", None)
