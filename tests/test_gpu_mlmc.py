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
