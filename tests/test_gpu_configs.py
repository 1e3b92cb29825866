"""GPU: statistical parity at the headline level and asserts on BASELINE.json's configs.

North star (b): "Philox-mode level means and variances agreeing with the reference within 3 standard errors".
The reference side is the pinned C restatement at 10^7 samples per case (tests/golden/cport_moments.json,
written by oracle/make_cport_moments.py - the Python reference needs days for that), the GPU side 10^8
samples, seeds fixed, both samplers.  Cases: l=6, M=4 (4096 fine steps, the benchmark level) for every GBM
product, l=7, M=2 (128 steps, block-uniform walk) and 16 / 4 steps (event-driven walk) for Merton."""
import json
import math
import os

import numpy as np
import pytest

import multilevel_mc_sdes_b200 as mm
from multilevel_mc_sdes_b200 import _lib
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

MOMENTS = json.load(open(os.path.join(GOLDEN, "cport_moments.json")))


def _mv(s, n):
    s = np.asarray(s, dtype=np.float64)
    m_d, m_f = s[0] / n, s[2] / n
    return m_d, s[1] / n - m_d ** 2, m_f, s[3] / n - m_f ** 2


@pytest.mark.parametrize("sampler", ["f64bm", "f32bm"])
@pytest.mark.parametrize("case", sorted(MOMENTS))
def test_level_moments_within_3_standard_errors(cuda_lib, case, sampler):
    ref = MOMENTS[case]
    n_ref, l, M = int(ref["n"]), int(ref["l"]), int(ref["M"])
    n_gpu = 100_000_000
    opt = getattr(mm, ref["cls"])(seed=4242, sampler=sampler, scheme=ref["scheme"], verbose=False, **ref["kwargs"])
    s = opt.looper(n_gpu, l, M, False)
    m_d, v_d, m_f, v_f = _mv(s, n_gpu)
    rm_d, rv_d, rm_f, rv_f = _mv(ref["sums"], n_ref)
    z_d = (m_d - rm_d) / math.sqrt(v_d / n_gpu + rv_d / n_ref)
    z_f = (m_f - rm_f) / math.sqrt(v_f / n_gpu + rv_f / n_ref)
    assert abs(z_d) < 3 and abs(z_f) < 3, (case, sampler, z_d, z_f)
    # variances: s.e. of a sample variance = v*sqrt((kurtosis - 1)/n).  The payoffs are heavy-tailed (kurtosis in
    # the hundreds for the level differences, thousands for digital ones), so the kurtosis is measured on 2*10^6
    # per-path outputs of the same kernel rather than assumed (x1.5 for its own sampling error)
    Pf, Pc = opt.payoff(2_000_000, l, M)
    for v, rv, x in ((v_f, rv_f, Pf), (v_d, rv_d, Pf - Pc)):
        c = x - x.mean()
        kurt = float(np.mean(c ** 4) / np.mean(c ** 2) ** 2)
        se = math.sqrt(1.5 * max(kurt - 1.0, 2.0) * (1.0 / n_ref + 1.0 / n_gpu))
        assert abs(v / rv - 1) < 3 * se, (case, sampler, v, rv, kurt)
    assert s[1] == pytest.approx(s[3] - 2 * s[6] + s[5], rel=1e-9, abs=1e-9 * s[3])


def test_samplers_agree_at_the_headline_level(cuda_lib):
    """The two samplers against each other at l=6, M=4, 2*10^8 samples each: |z| < 3 on E[dP] and E[Pf]."""
    n = 200_000_000
    kw = dict(X0=125, K=105, r=0.02, sig=0.5, verbose=False)
    a = _mv(mm.Euro_GBM(seed=11, sampler="f64bm", **kw).looper(n, 6, 4, False), n)
    b = _mv(mm.Euro_GBM(seed=12, sampler="f32bm", **kw).looper(n, 6, 4, False), n)
    assert abs(a[0] - b[0]) < 3 * math.sqrt((a[1] + b[1]) / n)
    assert abs(a[2] - b[2]) < 3 * math.sqrt((a[3] + b[3]) / n)
    assert abs(a[1] / b[1] - 1) < 3 * math.sqrt(2 * 600 / n) and abs(a[3] / b[3] - 1) < 3 * math.sqrt(2 * 12 / n)


def test_config3_lookback_level_sweep_1e8_per_level(cuda_lib):
    """BASELINE config #3: Lookback_GBM, fixed levels l = 0..8, 10^8 samples per level (Giles_plot's sweep
    :1775-1777 with its default M = 2).  The telescoping sum is the level-8 estimator: it must sit within its
    own statistical error plus the remaining weak error (first-order: about |E[dP_8]|) of BS(); variances and
    means of dP must decay at the rates the reference's regression (:1818-1823) would print."""
    opt = mm.Lookback_GBM(verbose=False, seed=3)
    n, L, M = 100_000_000, 8, 2
    sums = opt.level_sweep(Lmax=L, Nsamples=n, M=M)
    assert sums.shape == (7, L + 1)
    Y = sums[0] / n
    V = sums[1] / n - Y ** 2
    est, se = Y.sum(), math.sqrt(V.sum() / n)
    assert abs(est - opt.BS()) < 3 * se + 2 * abs(Y[-1]), (est, opt.BS(), se, Y[-1])
    lv = np.arange(1, L + 1)
    beta = -np.polyfit(lv, np.log2(V[1:]), 1)[0]
    alpha = -np.polyfit(lv, np.log2(np.abs(Y[1:])), 1)[0]
    assert 0.8 < beta < 1.3 and 0.7 < alpha < 1.3, (alpha, beta)
    # l == 0 row (:949) and the 7-sum identity on every other level
    assert sums[0, 0] == sums[2, 0] and sums[1, 0] == sums[3, 0] and not sums[4:, 0].any()
    for l in range(1, L + 1):
        assert sums[1, l] == pytest.approx(sums[3, l] - 2 * sums[6, l] + sums[5, l], rel=1e-9, abs=1e-9 * sums[3, l])
    # E[Pc_l] = E[Pf_{l-1}]: the coupling uses the same discretisation one level down
    for l in range(1, L + 1):
        vc = sums[5, l] / n - (sums[4, l] / n) ** 2
        vf = sums[3, l - 1] / n - (sums[2, l - 1] / n) ** 2
        assert abs(sums[4, l] / n - sums[2, l - 1] / n) < 4 * math.sqrt((vc + vf) / n), l


def _digital_merton_exact(o, nmax=80):
    """K e^{-rT} sum_n Poisson(n; lam T) Phi(d2_n): conditional on n jumps ln S_T is normal with mean
    ln X0 + (r - lam J_bar - sig^2/2) T + n m and variance sig^2 T + n s^2 (Merton 1976)."""
    tot = 0.0
    for n in range(nmax):
        w = math.exp(-o.lam * o.T) * (o.lam * o.T) ** n / math.factorial(n)
        d2 = (math.log(o.X0 / o.K) + (o.r - o.lam * o.J_bar - 0.5 * o.sig ** 2) * o.T + n * o.jumpmean) \
            / math.sqrt(o.sig ** 2 * o.T + n * o.jumpstd ** 2)
        tot += w * 0.5 * math.erfc(-d2 / math.sqrt(2.0))
    return o.K * math.exp(-o.r * o.T) * tot


@pytest.mark.parametrize("sampler", ["f64bm", "f32bm"])
def test_config4_digital_merton_mlmc(cuda_lib, sampler):
    """BASELINE config #4: Digital_Merton.mlmc(eps), M = 2, beta_0 = 0.5 (digital payoffs: variance decays like
    h^(1/2)).  The reference has no BS() for it; the Poisson-mixture closed form is the known answer."""
    eps = 2e-3
    errs = []
    for seed in (0, 1):
        opt = mm.Digital_Merton(beta_0=0.5, verbose=False, seed=seed, sampler=sampler)
        sums, N = opt.mlmc(eps, M=2)
        errs.append(float(np.sum(sums[0] / N)) - _digital_merton_exact(opt))
        assert len(N) >= 6 and N[0] > N[-1]
    assert max(abs(e) for e in errs) < 3 * eps, errs


def test_config2_asian_mlmc_tight(cuda_lib):
    """BASELINE config #2 one decade up (eps = 1e-3, M = 4, alpha_0 = beta_0 = None): against the value the
    reference driver converges to (Asian_GBM.BS() is Turnbull-Wakeman's approximation, 5.7828,
    not the limit)."""
    opt = mm.Asian_GBM(verbose=False, seed=5)
    sums, N = opt.mlmc(1e-3, M=4)
    assert abs(float(np.sum(sums[0] / N)) - 5.7630) < 3e-3


def test_pipe_and_mix_probes(cuda_lib):
    """The probes behind the roofline line (bench.py `roofline.pipe_probe`): every mode runs, the measured
    issue intervals are the ones the model rests on - IMAD.WIDE and LOP3 as long as DFMA (two SMSP cycles per warp
    instruction), FFMA half, IMAD.HI twice that, DFMA+LOP3 overlapping and DFMA+IMAD.WIDE additive - and the loop's
    mix finely interleaved beats the same mix on specialised warps."""
    import ctypes as C

    def interval(which):        # seconds per warp instruction per GPU: only ratios are used (clock-independent)
        ips, ms = C.c_double(), C.c_float()
        _lib.check(cuda_lib.mlmcb_pipe_probe(which, 0, C.byref(ips), C.byref(ms)))
        return 1.0 / ips.value

    dfma, ffma, wide, lop3, hi = interval(0), interval(2), interval(3), interval(4), interval(15)
    assert 0.8 < wide / dfma < 1.25 and 0.8 < lop3 / dfma < 1.25 and 0.4 < ffma / dfma < 0.65 and 1.6 < hi / dfma < 2.4
    pair_lop3, pair_wide = 2 * interval(9), 2 * interval(10)
    assert pair_lop3 < dfma + 0.6 * lop3          # mostly overlapped
    assert pair_wide > 0.85 * (dfma + wide)       # additive: IMAD.WIDE shares the FP64 pipe's port
    with pytest.raises(ValueError):
        _lib.check(cuda_lib.mlmcb_pipe_probe(99, 0, None, None))
    t = []
    for mode in (0, 1):
        ms = C.c_float()
        _lib.check(cuda_lib.mlmcb_mix_probe(mode, 0, 20000, C.byref(ms)))
        t.append(ms.value)
    assert 0 < t[0] < t[1]
