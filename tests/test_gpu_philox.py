"""GPU: Philox-mode level function through the C ABI (north-star check (b)) and its invariants."""
import ctypes as C

import numpy as np
import pytest

import multilevel_mc_sdes_b200 as mm
from multilevel_mc_sdes_b200 import _lib
from oracle import np_oracle as O
from tests.helpers import CLASS_CODES

pytestmark = pytest.mark.gpu


def level_sums(opt, n, l, M, offset=0, flags=None, per_path=False):
    import torch
    dev = torch.device("cuda", 0)
    out = torch.zeros(7, dtype=torch.float64, device=dev)
    Pf = torch.empty(n if per_path else 1, dtype=torch.float64, device=dev)
    Pc = torch.empty(n if per_path else 1, dtype=torch.float64, device=dev)
    p = opt._params()
    _lib.check(_lib.load().mlmcb_level_sums(
        C.byref(p), opt._model, opt._payoff_kind, l, M, n, opt.seed, offset,
        opt._flags() if flags is None else flags, 0, 0, out.data_ptr(),
        Pf.data_ptr() if per_path else None, Pc.data_ptr() if per_path else None))
    torch.cuda.synchronize()
    if per_path:
        return out.cpu().numpy(), Pf.cpu().numpy(), Pc.cpu().numpy()
    return out.cpu().numpy()


@pytest.mark.parametrize("sampler", [_lib.SAMPLER_F32BM, _lib.SAMPLER_F64BM])
def test_sampler_is_standard_normal(cuda_lib, sampler):
    import torch
    from scipy import stats
    n_paths, n_steps = 1 << 16, 64
    out = torch.empty(n_steps * n_paths, dtype=torch.float64, device="cuda")
    _lib.check(cuda_lib.mlmcb_sampler_probe(sampler, 12345, 3, 0, n_paths, n_steps, 0, 0, out.data_ptr()))
    torch.cuda.synchronize()
    z = out.cpu().numpy()
    N = z.size                                  # 4.2e6 draws
    assert abs(z.mean()) < 4 / np.sqrt(N)
    assert abs(z.var() - 1) < 4 * np.sqrt(2 / N)
    assert abs(stats.skew(z)) < 4 * np.sqrt(6 / N)
    assert abs(stats.kurtosis(z)) < 4 * np.sqrt(24 / N)
    assert stats.kstest(z[:1000000], "norm").pvalue > 1e-3
    assert 4.5 < np.abs(z).max() < 7.5          # tails are populated and finite
    # tail mass: P(|z|>3) = 2.6998e-3
    p3 = (np.abs(z) > 3).mean()
    assert abs(p3 - 2.6998e-3) < 4 * np.sqrt(2.6998e-3 / N)
    # consecutive steps of a path and neighbouring paths are uncorrelated
    zz = z.reshape(n_steps, n_paths)
    assert abs(np.mean(zz[:-1] * zz[1:])) < 4 / np.sqrt(zz[:-1].size)
    assert abs(np.mean(zz[:, :-1] * zz[:, 1:])) < 4 / np.sqrt(zz[:, :-1].size)
    # a different level tag / seed / path offset gives a different stream; same arguments the same one
    out2 = torch.empty_like(out)
    _lib.check(cuda_lib.mlmcb_sampler_probe(sampler, 12345, 3, 0, n_paths, n_steps, 0, 0, out2.data_ptr()))
    assert torch.equal(out, out2)
    _lib.check(cuda_lib.mlmcb_sampler_probe(sampler, 12345, 4, 0, n_paths, n_steps, 0, 0, out2.data_ptr()))
    assert abs(np.corrcoef(z, out2.cpu().numpy())[0, 1]) < 4 / np.sqrt(N)
    _lib.check(cuda_lib.mlmcb_sampler_probe(sampler, 12345, 3, 1000, n_paths, n_steps, 0, 0, out2.data_ptr()))
    z3 = out2.cpu().numpy().reshape(n_steps, n_paths)
    assert np.array_equal(z3[:, :-1000], zz[:, 1000:])       # ids are global: offset shifts the window


def _moments(s, n, l):
    m_d, m_f = s[0] / n, s[2] / n
    v_d, v_f = s[1] / n - m_d ** 2, s[3] / n - m_f ** 2
    return m_d, v_d, m_f, v_f


CASES = [(c, l, M) for c in sorted(CLASS_CODES) for (l, M) in ((0, 2), (1, 2), (2, 4), (3, 4), (4, 2))]
# Merton above 128 fine steps runs the block-uniform walk (jump records in shared memory)
CASES += [(c, l, M) for c in sorted(CLASS_CODES) if CLASS_CODES[c][0] == O.MERTON for (l, M) in ((8, 2), (4, 4))]


@pytest.mark.parametrize("cname,l,M", CASES)
def test_level_moments_match_reference_port(cuda_lib, c_oracle, cname, l, M):
    """Means and variances of dP and Pf within 4 combined standard errors of the pinned C port of
    the reference (own generator), both samplers; plus the 7-sum identity on the GPU result."""
    model, pay = CLASS_CODES[cname]
    n_gpu, n_cpu = 4_000_000, 400_000
    ref = c_oracle.level_sums(O.Params(), model, pay, l, M, n_cpu, seed=11 + l, nthreads=8)
    rm_d, rv_d, rm_f, rv_f = _moments(ref, n_cpu, l)
    for sampler in ("f32bm", "f64bm"):
        opt = getattr(mm, cname)(seed=2024, sampler=sampler)
        s = level_sums(opt, n_gpu, l, M)
        m_d, v_d, m_f, v_f = _moments(s, n_gpu, l)
        se = lambda v: np.sqrt(v / n_gpu + v / n_cpu)
        assert abs(m_d - rm_d) < 4 * se(max(v_d, rv_d)) + 1e-15, (sampler, m_d, rm_d)
        assert abs(m_f - rm_f) < 4 * se(max(v_f, rv_f)), (sampler, m_f, rm_f)
        # variances: s.e. of a sample variance ~ sqrt((m4 - v^2)/n); bound by a generous kurtosis of 60
        assert abs(v_f - rv_f) < 4 * rv_f * np.sqrt(60 / n_cpu), (sampler, v_f, rv_f)
        if l > 0 and rv_d > 0:
            tol = 4 * rv_d * np.sqrt((3000 if pay == O.DIGITAL else 300) / n_cpu)
            assert abs(v_d - rv_d) < tol, (sampler, v_d, rv_d)
        if l == 0:
            assert s[0] == s[2] and s[1] == s[3] and not s[4:].any()
        else:
            assert s[1] == pytest.approx(s[3] - 2 * s[6] + s[5], rel=1e-9, abs=1e-9 * s[3])


def test_per_path_outputs_consistent_and_deterministic(cuda_lib):
    for cname in ("Euro_GBM", "Asian_GBM", "Lookback_GBM", "Digital_GBM", "Euro_Merton", "Asian_Merton",
                  "Lookback_Merton", "Digital_Merton"):
        opt = getattr(mm, cname)(seed=5)
        n = 100_003
        s, Pf, Pc = level_sums(opt, n, 3, 4, per_path=True)
        want = O.seven_sums(Pf, Pc, 3)
        np.testing.assert_allclose(s, want, rtol=1e-12)
        s2, Pf2, Pc2 = level_sums(opt, n, 3, 4, per_path=True)
        assert np.array_equal(s, s2) and np.array_equal(Pf, Pf2) and np.array_equal(Pc, Pc2)   # run-to-run
        # global ids: two half calls reproduce the per-path values of one call
        _, Pfa, _ = level_sums(opt, 50_000, 3, 4, offset=0, per_path=True)
        _, Pfb, _ = level_sums(opt, n - 50_000, 3, 4, offset=50_000, per_path=True)
        assert np.array_equal(np.concatenate([Pfa, Pfb]), Pf)
        # another seed gives other paths
        s3 = level_sums(getattr(mm, cname)(seed=6), n, 3, 4)
        assert s3[2] != s[2]


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
def test_level0_mean_against_closed_form(cuda_lib, sampler):
    """l = 0 is one Euler step, X = X0(1+rT) + X0 sig sqrt(T) z, so E[Pf] is known exactly (normal X):
    4e9 samples resolve 1e-5 in the mean of z.  (tools/l0_exact_check.py runs this at 3e12 samples.)"""
    import math
    Phi = lambda x: 0.5 * math.erfc(-x / math.sqrt(2.0))
    phi = lambda x: math.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    n = 4_000_000_000 if sampler == "f32bm" else 1_000_000_000
    for kw in ({}, dict(X0=125, K=105, r=0.02, sig=0.5)):
        o = mm.Euro_GBM(seed=321, sampler=sampler, **kw)
        m, s_, disc = o.X0 * (1 + o.r * o.T), o.X0 * o.sig * math.sqrt(o.T), math.exp(-o.r * o.T)
        d = (m - o.K) / s_
        exact = disc * ((m - o.K) * Phi(d) + s_ * phi(d))
        second = disc ** 2 * (((m - o.K) ** 2 + s_ * s_) * Phi(d) + (m - o.K) * s_ * phi(d))
        sums = o.looper(n, 0, 2, False)
        assert abs(sums[2] / n - exact) < 4 * math.sqrt((second - exact ** 2) / n)
        assert abs(sums[3] / n - second) < 4 * second * math.sqrt(30 / n)          # second moment, loose kurtosis bound


def test_merton_walks_ragged_sizes(cuda_lib):
    """Fewer paths than a warp / a block, sizes that are no multiple of anything, on each Merton walk
    (1 step: lane-queue loop, 4 and 64: event-driven, 256: block-uniform): sums equal the per-path outputs,
    and a path's payoff does not depend on how many neighbours were simulated with it."""
    for cls in (mm.Euro_Merton, mm.Asian_Merton, mm.Lookback_Merton):
        for l, M in ((0, 2), (2, 2), (3, 4), (8, 2)):
            ref = cls(seed=8)
            ref._next_path[l] = 40
            Rf, Rc = ref.payoff(300, l, M)
            for n in (1, 5, 33, 129):
                a = cls(seed=8)
                a._next_path[l] = 40
                Pf, Pc = a.payoff(n, l, M)
                assert np.array_equal(Pf, Rf[:n]) and np.array_equal(Pc, Rc[:n]), (cls.__name__, l, M, n)
                s = level_sums(a, n, l, M, offset=40)
                want = [Pf.sum(), (Pf ** 2).sum(), Pf.sum(), (Pf ** 2).sum(), 0, 0, 0] if l == 0 else \
                    [(Pf - Pc).sum(), ((Pf - Pc) ** 2).sum(), Pf.sum(), (Pf ** 2).sum(), Pc.sum(), (Pc ** 2).sum(), (Pf * Pc).sum()]
                np.testing.assert_allclose(s, want, rtol=1e-12, atol=1e-12)


def test_coupling_fine_equals_previous_level_distribution(cuda_lib):
    """Telescoping check: E[Pc at level l] = E[Pf at level l-1] (same discretisation)."""
    for cname in ("Euro_GBM", "Asian_GBM", "Lookback_GBM", "Euro_Merton", "Asian_Merton", "Lookback_Merton"):
        n = 4_000_000
        a = level_sums(getattr(mm, cname)(seed=1), n, 3, 4)
        b = level_sums(getattr(mm, cname)(seed=2), n, 2, 4)
        mc, vc = a[4] / n, a[5] / n - (a[4] / n) ** 2
        mf, vf = b[2] / n, b[3] / n - (b[2] / n) ** 2
        assert abs(mc - mf) < 4 * np.sqrt(vc / n + vf / n), cname


def test_generic_M_and_edge_sizes(cuda_lib, c_oracle):
    # odd M takes the generic (counter) path; compare with the C port
    for cname in ("Euro_GBM", "Asian_Merton"):
        model, pay = CLASS_CODES[cname]
        n = 2_000_000
        s = level_sums(getattr(mm, cname)(seed=3), n, 3, 3)
        ref = c_oracle.level_sums(O.Params(), model, pay, 3, 3, 300_000, seed=4, nthreads=8)
        m, v = s[0] / n, s[1] / n - (s[0] / n) ** 2
        rm = ref[0] / 300_000
        assert abs(m - rm) < 4 * np.sqrt(v / n + v / 300_000)
    opt = mm.Euro_GBM()
    assert not level_sums(opt, 0, 2, 2).any()                      # empty request
    s1 = level_sums(opt, 1, 2, 2)                                  # a single path
    assert s1[3] == pytest.approx(s1[2] ** 2)
    big = level_sums(opt, 1000, 10, 4)                             # 1 048 576 steps per path
    assert np.isfinite(big).all()


def test_fp32_opt_in_runs_and_is_close(cuda_lib):
    n = 2_000_000
    opt64, opt32 = mm.Asian_GBM(seed=9), mm.Asian_GBM(seed=9, fp32=True)
    a, b = level_sums(opt64, n, 3, 4), level_sums(opt32, n, 3, 4)
    assert a[2] / n == pytest.approx(b[2] / n, rel=1e-4)
    assert a[0] / n == pytest.approx(b[0] / n, abs=5e-3)


def test_looper_host_call_and_sweep(cuda_lib):
    opt = mm.Lookback_GBM(seed=77)
    s = opt.looper(200_000, 2, 4, False)
    assert s.shape == (7,) and s.dtype == np.float64
    opt2 = mm.Lookback_GBM(seed=77)
    sw = opt2.level_sweep(Lmax=4, Nsamples=200_000, M=4)
    assert sw.shape == (7, 5)
    # same ids, same paths, sweep or single call; the sweep kernel cuts the paths into other block ranges, so
    # the sums agree to summation order
    np.testing.assert_allclose(sw[:, 2], s, rtol=1e-13)
    # second call on the same object uses fresh ids
    assert not np.array_equal(opt.looper(200_000, 2, 4, False), s)
    Pf, Pc = mm.Euro_GBM().payoff(1000, 0, 2)
    assert Pf.shape == (1000,) and np.all(Pc == 100.0)             # l==0: raw coarse output (:82)


@pytest.mark.parametrize("cname,l,M", [(c, l, M) for c in ("Euro_GBM", "Asian_GBM") for (l, M) in ((0, 2), (1, 2), (2, 4), (3, 4), (2, 6))])
def test_antithetic_level_moments(cuda_lib, c_oracle, cname, l, M):
    """looper(anti=True): moments vs the pinned C port; the antithetic estimator keeps the mean of dP
    (with plain Euler-Maruyama it does not reduce the variance order, so nothing is claimed on that)."""
    _, pay = CLASS_CODES[cname]
    n_gpu, n_cpu = 4_000_000, 400_000
    ref = c_oracle.level_sums_anti(O.Params(), pay, l, M, n_cpu, seed=21 + l, nthreads=8)
    opt = getattr(mm, cname)(seed=31)
    s = opt.looper(n_gpu, l, M, True)
    m_d, v_d, m_f, v_f = _moments(s, n_gpu, l)
    rm_d, rv_d, rm_f, rv_f = _moments(ref, n_cpu, l)
    se = lambda v: np.sqrt(v / n_gpu + v / n_cpu)
    assert abs(m_d - rm_d) < 4 * se(max(v_d, rv_d)) + 1e-15
    assert abs(m_f - rm_f) < 4 * se(max(v_f, rv_f))
    if l > 0:
        assert abs(v_d - rv_d) < 4 * rv_d * np.sqrt(300 / n_cpu)
        plain = getattr(mm, cname)(seed=32).looper(n_gpu, l, M, False)
        pm, pv, _, _ = _moments(plain, n_gpu, l)
        assert abs(m_d - pm) < 4 * np.sqrt(v_d / n_gpu + pv / n_gpu)       # same expectation
        assert s[1] == pytest.approx(s[3] - 2 * s[6] + s[5], rel=1e-9, abs=1e-9 * s[3])
    Pf, Pc = opt.anti_payoff(1000, l, M)
    assert Pf.shape == (1000,) and Pc.shape == (1000,)
    if l == 0:      # no antithetic twin at l = 0 (:212-214, :310): the sums are the plain kernel's
        a, b = getattr(mm, cname)(seed=77), getattr(mm, cname)(seed=77)
        assert np.array_equal(a.looper(123_457, 0, M, True), b.looper(123_457, 0, M, False))
        assert np.all(Pc == opt.X0)


@pytest.mark.parametrize("cname", ["Euro_GBM", "Asian_GBM", "Lookback_GBM", "Euro_Merton", "Asian_Merton"])
def test_milstein_level_moments(cuda_lib, c_oracle, cname):
    model, pay = CLASS_CODES[cname]
    n_gpu, n_cpu = 4_000_000, 400_000
    for (l, M) in ((0, 2), (2, 4), (3, 2)):
        ref = c_oracle.level_sums(O.Params(scheme="milstein"), model, pay, l, M, n_cpu, seed=5 + l, nthreads=8)
        s = getattr(mm, cname)(seed=77, scheme="milstein").looper(n_gpu, l, M, False)
        m_d, v_d, m_f, v_f = _moments(s, n_gpu, l)
        rm_d, rv_d, rm_f, rv_f = _moments(ref, n_cpu, l)
        se = lambda v: np.sqrt(v / n_gpu + v / n_cpu)
        assert abs(m_d - rm_d) < 4 * se(max(v_d, rv_d)) + 1e-15, (l, M, m_d, rm_d)
        assert abs(m_f - rm_f) < 4 * se(max(v_f, rv_f))
        if l > 0:
            assert abs(v_d - rv_d) < 4 * rv_d * np.sqrt(300 / n_cpu)
    # Milstein has strong order 1: the variance of dP falls ~M^2 per level for the European payoff
    if cname == "Euro_GBM":
        o = mm.Euro_GBM(seed=1, scheme="milstein")
        v = []
        for l in (2, 3, 4):
            s = o.looper(2_000_000, l, 4, False)
            v.append(s[1] / 2e6 - (s[0] / 2e6) ** 2)
        assert 10 < v[0] / v[1] < 24 and 10 < v[1] / v[2] < 24
    with pytest.raises(NotImplementedError):
        mm.Euro_GBM(scheme="milstein", fp32=True).looper(100, 2, 2, False)      # Milstein: fp64 path arithmetic only


def test_f64bm_stream_matches_float64_recomputation(cuda_lib):
    """The fp64 sampler's normals recomputed on the host from the same Philox4x32-10 words (counter layout of
    mlmcb_kernels.cuh: c0 = stream, c1 = Philox block index, c2,c3 = path id | level<<16; steps 8o+1..8o+8 take
    blocks 3o..3o+2, three words per pair) with numpy float64 log/sqrt/sin/cos: the table-driven device math
    agrees to < 1e-14."""
    import torch
    seed, level, n_paths, n_steps, off = 987654321012345, 5, 64, 16, 7_000_000_123
    out = torch.empty(n_steps * n_paths, dtype=torch.float64, device="cuda")
    _lib.check(cuda_lib.mlmcb_sampler_probe(_lib.SAMPLER_F64BM, seed, level, off, n_paths, n_steps, 0, 0, out.data_ptr()))
    torch.cuda.synchronize()
    z = out.cpu().numpy().reshape(n_steps, n_paths)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    worst = 0.0
    for p in range(n_paths):
        pid = off + p
        for o in range(n_steps // 8):
            w = []
            for blk in range(3):
                w += _lib.philox4x32_10((0, 3 * o + blk, pid & 0xFFFFFFFF, ((pid >> 32) & 0xFFFF) | (level << 16)), key)
            for q in range(4):
                a, b, c = w[3 * q:3 * q + 3]
                u = (((a & 0xFFFFF) << 32 | b) + 0.5) * 2.0 ** -52          # 52 random bits + half-ulp offset
                h = (((a >> 23) & 0xFF) + c * 2.0 ** -32) / 256.0 - 0.5      # 8 + 32 angle bits
                rad = np.sqrt(-2.0 * np.log(u)) * (-1.0 if a >> 31 else 1.0)
                want = (rad * np.cos(np.pi * h), rad * np.sin(np.pi * h))
                got = (z[8 * o + 2 * q, p], z[8 * o + 2 * q + 1, p])
                worst = max(worst, abs(got[0] - want[0]), abs(got[1] - want[1]))
    assert worst < 1e-14, worst


def test_american_put_level_moments_and_price(cuda_lib, c_oracle):
    n_gpu, n_cpu = 4_000_000, 300_000
    for l in (0, 1, 3, 5):
        ref = c_oracle.level_sums(O.Params(), O.GBM, O.AMERICAN, l, 2, n_cpu, seed=40 + l, nthreads=8)
        for sampler in ("f32bm", "f64bm"):
            s = mm.Amer_GBM(seed=8, sampler=sampler).looper(n_gpu, l, 2, False)
            m_d, v_d, m_f, v_f = _moments(s, n_gpu, l)
            rm_d, rv_d, rm_f, rv_f = _moments(ref, n_cpu, l)
            se = lambda v: np.sqrt(v / n_gpu + v / n_cpu)
            assert abs(m_d - rm_d) < 4 * se(max(v_d, rv_d)) + 1e-15, (l, sampler)
            assert abs(m_f - rm_f) < 4 * se(max(v_f, rv_f)), (l, sampler)
            if l > 0:
                assert abs(v_d - rv_d) < 4 * rv_d * np.sqrt(300 / n_cpu)
    # the early-exercise premium: American put above the European put (BS put = call - S + K e^{-rT} = 5.5735)
    opt = mm.Amer_GBM(verbose=False, alpha_0=1, beta_0=1)
    sums, N = opt.mlmc(0.01, M=2)
    price = float(np.sum(sums[0] / N))
    assert 5.6 < price < 6.2, price          # reference boundary c=0.8 is sub-optimal; true value 6.09
    with pytest.raises(ValueError, match="M must be equal to 2"):
        opt.looper(100, 2, 4, False)
    with pytest.raises(NotImplementedError):
        mm.Amer_GBM(scheme="milstein").looper(100, 2, 2, False)


def test_path_functions_return_reference_shapes(cuda_lib):
    """Option.path / anti_path: the per-path quantities behind the payoffs (same ids -> same paths)."""
    disc = np.exp(-0.05)
    o = mm.Euro_GBM(seed=4)
    Xf, Xc = o.path(5000, 3, 4)
    o2 = mm.Euro_GBM(seed=4)
    Pf, Pc = o2.payoff(5000, 3, 4)
    assert np.array_equal(Pf, disc * np.maximum(0, Xf - 100)) and np.array_equal(Pc, disc * np.maximum(0, Xc - 100))
    Xf, Mf, Xc, Mc = mm.Lookback_GBM(seed=4).path(5000, 3, 4)
    assert np.all(Mf <= Xf) and np.all(Mc <= Xc) and np.all(Mf <= 100) and Xf.shape == (5000,)
    Af, Ac = mm.Asian_Merton(seed=4).path(2000, 2, 2)
    assert Af.shape == (2000,) and np.all(Af > 0) and np.all(Ac > 0)
    Xf, Xa, Xc = mm.Euro_GBM(seed=4).anti_path(3000, 3, 4)
    assert not np.array_equal(Xf, Xa) and abs(np.mean(Xf) - np.mean(Xa)) < 1.0
    Xf, Xa, Xc = mm.Asian_GBM(seed=4).anti_path(3000, 0, 4)
    assert np.array_equal(Xf, Xa)                                     # l==0: "just ignore Xa=Xf" (:215)
    F, Cc = mm.Amer_GBM(seed=4).path(4000, 4, 2)
    assert F.shape == (4000, 3) and Cc.shape == (4000, 3)
    assert np.all((F[:, 1] >= 0) & (F[:, 1] <= 1)) and np.all(F[:, 2] >= 0)
    with pytest.raises(NotImplementedError):
        mm.Option().path(10, 1, 2)
    with pytest.raises(NotImplementedError):
        mm.Digital_Merton().anti_path(10, 1, 2)


def test_maximum_sizes(cuda_lib):
    """Path ids beyond 2^32 (the id is 48 bits wide), more paths per launch than 2^32, and the longest
    supported path (2^20 fine steps)."""
    opt = mm.Euro_GBM(seed=3)
    n = 5_000_000_000                                               # > 2^32 paths in one launch, l = 0
    s = level_sums(opt, n, 0, 2)
    bs_l0_tol = 4 * np.sqrt((s[1] / n - (s[0] / n) ** 2) / n)
    ref = level_sums(opt, 50_000_000, 0, 2, offset=(1 << 40))      # ids around 2^40: same distribution
    assert abs(s[0] / n - ref[0] / 5e7) < bs_l0_tol + 4 * np.sqrt((ref[1] / 5e7 - (ref[0] / 5e7) ** 2) / 5e7)
    a = level_sums(opt, 1000, 2, 2, offset=(1 << 32) - 500, per_path=True)[1]
    b = level_sums(opt, 500, 2, 2, offset=(1 << 32), per_path=True)[1]
    assert np.array_equal(a[500:], b)                                # ids straddling 2^32 are consistent
    with pytest.raises(ValueError):
        level_sums(opt, 10, 2, 2, offset=(1 << 48))                  # ids must stay below 2^48
    long = level_sums(mm.Asian_GBM(seed=1), 20_000, 20, 2)           # 1 048 576 fine steps per path
    assert np.isfinite(long).all() and long[3] > 0
    with pytest.raises(ValueError):
        level_sums(opt, 10, 34, 2)                                   # 2^34 steps: above the 2^33 cap


def test_concurrent_host_calls_from_threads(cuda_lib):
    """The _host entry points are serialised per device and the workspace is stream-ordered: calls from
    several Python threads give the same sums as the same calls made one after the other."""
    import threading
    specs = [(mm.Euro_GBM, 3, 4), (mm.Asian_GBM, 2, 2), (mm.Lookback_Merton, 3, 2), (mm.Digital_GBM, 4, 2),
             (mm.Amer_GBM, 3, 2), (mm.Euro_Merton, 1, 2)]
    want = [cls(seed=9).looper(200_003, l, M, False) for cls, l, M in specs]
    got = [None] * len(specs)

    def work(i):
        cls, l, M = specs[i]
        for _ in range(3):
            got[i] = cls(seed=9).looper(200_003, l, M, False)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(specs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_looper_async_matches_looper(cuda_lib):
    import torch
    a = mm.Asian_GBM(seed=12).looper(300_000, 3, 4, False)
    t = mm.Asian_GBM(seed=12).looper_async(300_000, 3, 4)
    assert t.is_cuda and t.dtype == torch.float64
    assert np.array_equal(t.cpu().numpy(), a)


def test_packed_low_levels_are_id_addressed(cuda_lib, c_oracle):
    """l = 0 (and M=2, l=1) share one Philox block between 4 (2) neighbouring path ids: the per-path
    values must not depend on how a range is cut (odd offsets, odd lengths), only on the global id."""
    for cls, l, M in ((mm.Euro_GBM, 0, 2), (mm.Asian_GBM, 0, 4), (mm.Lookback_GBM, 1, 2), (mm.Digital_GBM, 0, 3),
                      (mm.Euro_GBM, 1, 2)):
        opt = cls(seed=21)
        _, Pf, Pc = level_sums(opt, 10_007, l, M, offset=3, per_path=True)
        _, Pa, _ = level_sums(opt, 4_001, l, M, offset=3, per_path=True)
        _, Pb, Pcb = level_sums(opt, 6_006, l, M, offset=4_004, per_path=True)
        assert np.array_equal(np.concatenate([Pa, Pb]), Pf)
        assert np.array_equal(Pcb, Pc[4_001:])
        s = level_sums(opt, 10_007, l, M, offset=3)
        np.testing.assert_allclose(s, O.seven_sums(Pf, Pc, l), rtol=1e-12)
        assert len(np.unique(Pf[Pf > 0])) > 0.4 * np.count_nonzero(Pf > 0) or cls is mm.Digital_GBM   # distinct draws
    # neighbours inside one block are independent draws
    _, Pf, _ = level_sums(mm.Lookback_GBM(seed=2), 400_000, 0, 2, per_path=True)
    q = Pf.reshape(-1, 4)
    c = np.corrcoef(q.T)
    assert np.all(np.abs(c[np.triu_indices(4, 1)]) < 4 / np.sqrt(len(q)))


def _probe_normals(cuda_lib, sampler, seed, l, offset, n_paths, n_steps):
    import torch
    out = torch.empty(n_steps * n_paths, dtype=torch.float64, device="cuda")
    _lib.check(cuda_lib.mlmcb_sampler_probe(sampler, seed, l, offset, n_paths, n_steps, 0, 0, out.data_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(n_steps, n_paths)


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
def test_philox_mode_equals_replay_of_its_own_draws_gbm(cuda_lib, sampler):
    """Deterministic cross-check of the production kernels: the normals a Philox-mode call uses
    (sampler probe) pushed through the replay entry point reproduce its per-path payoffs."""
    code = _lib.SAMPLER_F32BM if sampler == "f32bm" else _lib.SAMPLER_F64BM
    for cls, scheme in ((mm.Euro_GBM, "em"), (mm.Asian_GBM, "em"), (mm.Lookback_GBM, "em"), (mm.Digital_GBM, "em"),
                        (mm.Asian_GBM, "milstein"), (mm.Lookback_GBM, "milstein")):
        # (1,4) and (2,2): the four-step kernels; (6,4): the headline level (4096 steps)
        for l, M in ((2, 2), (1, 4), (3, 4), (2, 3), (5, 2), (6, 4)):
            n, off, seed = 1000, 12345, 31
            opt = cls(seed=seed, sampler=sampler, scheme=scheme)
            opt._next_path[l] = off
            Pf, Pc = opt.payoff(n, l, M)
            Z = _probe_normals(cuda_lib, code, seed, l, off, n, M ** l)
            Rf, Rc = cls(scheme=scheme).payoff_replay(Z, l, M, arith="fast")
            np.testing.assert_allclose(Pf, Rf, rtol=1e-13, atol=1e-11)
            np.testing.assert_allclose(Pc, Rc, rtol=1e-13, atol=1e-11)
    # American put (its own path kernel, same grid-step stream)
    for l in (0, 1, 4, 7):
        opt = mm.Amer_GBM(seed=6, sampler=sampler)
        opt._next_path[l] = 41
        Pf, Pc = opt.payoff(700, l, 2)
        Z = _probe_normals(cuda_lib, code, 6, l, 41, 700, 2 ** l)
        Rf, Rc = mm.Amer_GBM().payoff_replay(Z, l, 2, arith="fast")
        np.testing.assert_allclose(Pf, Rf, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(Pc, Rc, rtol=1e-12, atol=1e-10)
    # antithetic, with either time-stepper
    for scheme in ("em", "milstein"):
        for cls in (mm.Asian_GBM, mm.Euro_GBM):
            opt = cls(seed=5, sampler=sampler, scheme=scheme)
            opt._next_path[3] = 77
            Pf, Pc = opt.anti_payoff(500, 3, 4)
            Z = _probe_normals(cuda_lib, code, 5, 3, 77, 500, 64)
            Rf, Rc = cls(scheme=scheme).payoff_replay(Z, 3, 4, arith="fast", anti=True)
            np.testing.assert_allclose(Pf, Rf, rtol=1e-13, atol=1e-11)
            np.testing.assert_allclose(Pc, Rc, rtol=1e-13, atol=1e-11)


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
def test_philox_mode_equals_replay_of_its_own_draws_gbm_random_parameters(cuda_lib, sampler):
    """The same with random model parameters, path-id offsets and levels: the packed kernel (1-2 steps), the
    level kernel, and the American-put kernel with a host-tabulated boundary (T != 1 moves every table time)."""
    import torch
    from oracle.boundaries import BOUNDARIES
    code = _lib.SAMPLER_F32BM if sampler == "f32bm" else _lib.SAMPLER_F64BM
    rng = np.random.default_rng(31415)
    classes = (mm.Euro_GBM, mm.Asian_GBM, mm.Lookback_GBM, mm.Digital_GBM)
    levels = ((0, 2), (1, 2), (0, 4), (2, 2), (1, 4), (1, 3), (3, 2), (2, 4), (2, 5), (6, 2), (3, 4))
    for it in range(22):
        kw = dict(X0=float(rng.uniform(5, 400)), T=float(rng.uniform(0.1, 4.0)), r=float(rng.uniform(-0.01, 0.12)),
                  sig=float(rng.uniform(0.02, 0.8)))
        kw["K"] = kw["X0"] * float(rng.uniform(0.6, 1.4))
        cls = classes[it % 4]
        l, M = levels[it % len(levels)]
        n, off, seed = 500, int(rng.integers(0, 2 ** 44)), int(rng.integers(0, 2 ** 31))
        opt = cls(seed=seed, sampler=sampler, **kw)
        opt._next_path[l] = off
        Pf, Pc = opt.payoff(n, l, M)
        nsteps = M ** l
        if nsteps <= 2:
            out = torch.empty(nsteps * n, dtype=torch.float64, device="cuda")
            _lib.check(cuda_lib.mlmcb_packed_probe(code, seed, l, nsteps, off, n, 0, 0, out.data_ptr()))
            torch.cuda.synchronize()
            Z = out.cpu().numpy().reshape(nsteps, n)
        else:
            Z = _probe_normals(cuda_lib, code, seed, l, off, n, nsteps)
        Rf, Rc = cls(**kw).payoff_replay(Z, l, M, arith="fast")
        scale = 1e-12 * max(kw["X0"], kw["K"])
        np.testing.assert_allclose(Pf, Rf, rtol=1e-12, atol=scale)
        np.testing.assert_allclose(Pc, Rc, rtol=1e-12, atol=scale)
    for bname, l in (("power", 3), ("linear", 6), ("power", 0)):
        kw = dict(X0=float(rng.uniform(50, 150)), T=float(rng.uniform(0.3, 2.5)), r=0.04, sig=float(rng.uniform(0.1, 0.5)))
        kw["K"] = kw["X0"] * 1.05
        opt = mm.Amer_GBM(exerciseBoundary=BOUNDARIES[bname], seed=12, sampler=sampler, **kw)
        opt._next_path[l] = 9000
        Pf, Pc = opt.payoff(400, l, 2)
        Z = _probe_normals(cuda_lib, code, 12, l, 9000, 400, 2 ** l)
        Rf, Rc = mm.Amer_GBM(exerciseBoundary=BOUNDARIES[bname], **kw).payoff_replay(Z, l, 2, arith="fast")
        np.testing.assert_allclose(Pf, Rf, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(Pc, Rc, rtol=1e-12, atol=1e-10)


def _merton_own_draws_case(cuda_lib, cls, kw, l, M, n=600, off=999, seed=17, K=64, sampler="f64bm"):
    """Per-path payoffs of the production Merton kernel against the replay of its own draws: grid normals
    from the sampler probe, jump records from the jump probe, interleaved on the host in the order the
    reference consumes them (:454-479)."""
    import torch
    code = _lib.SAMPLER_F32BM if sampler == "f32bm" else _lib.SAMPLER_F64BM
    opt = cls(seed=seed, sampler=sampler, **kw)
    opt._next_path[l] = off
    Pf, Pc = opt.payoff(n, l, M)
    nsteps = M ** l
    # grid normals: the coarsest GBM-style packing does not apply to Merton; block-addressed stream
    Z = _probe_normals(cuda_lib, code, seed, l, off, n, max(nsteps, 1))
    rec = torch.empty(3 * K * n, dtype=torch.float64, device="cuda")
    _lib.check(cuda_lib.mlmcb_jump_probe(code, seed, l, opt.lam, off, n, K, 0, 0, rec.data_ptr()))
    torch.cuda.synchronize()
    rec = rec.cpu().numpy().reshape(3, K, n)
    zW, zQ, E, oW, oQ, oE = [], [], [], [0], [0], [0]
    dt = opt.T / nsteps
    for i in range(n):
        k = 0
        tau = rec[0, 0, i]
        E.append(rec[0, 0, i])
        for j in range(1, nsteps + 1):
            tn = j * dt                     # the production kernel's grid time (double)j*dt
            while tau < tn:
                zW.append(rec[2, k, i]); zQ.append(rec[1, k, i])
                k += 1
                assert k < K
                E.append(rec[0, k, i])
                tau += rec[0, k, i]
            # l = 0: the closing step takes the normal of the record that overshoots T
            zW.append(rec[2, k, i] if nsteps == 1 else Z[j - 1, i])
        oW.append(len(zW)); oQ.append(len(zQ)); oE.append(len(E))
    packed = dict(zW=np.array(zW), offW=np.array(oW, dtype=np.int64), zQ=np.array(zQ),
                  offQ=np.array(oQ, dtype=np.int64), E=np.array(E), offE=np.array(oE, dtype=np.int64))
    Rf, Rc = cls(**kw).payoff_replay(packed, l, M, arith="fast")
    scale = 1e-12 * max(opt.X0, opt.K or 0.0)
    np.testing.assert_allclose(Pf, Rf, rtol=1e-12, atol=100 * scale)
    np.testing.assert_allclose(Pc, Rc, rtol=1e-12, atol=100 * scale)


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
@pytest.mark.parametrize("lam", [1.0, 6.0])
def test_philox_mode_equals_replay_of_its_own_draws_merton(cuda_lib, lam, sampler):
    """Exercises the jump-free run length agreement, the jump blocks and the tail of the block-uniform kernel,
    the event-driven kernel and the l = 0 loop against the straightforward replay loop."""
    for cls in (mm.Euro_Merton, mm.Asian_Merton, mm.Lookback_Merton):
        # 1 step: lane-queue loop; 9 (M = 3): event-driven walk; everything else: the factor walk - 2 steps in one go, 8 and
        # 32 step by step, 64 .. 512 with jump-free runs and (final-value products) mixed octets; lam = 6 needs record rounds
        for l, M in ((0, 2), (1, 2), (3, 2), (3, 4), (2, 3), (5, 2), (7, 2), (4, 4), (9, 2)):
            _merton_own_draws_case(cuda_lib, cls, dict(lam=lam), l, M, sampler=sampler)


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
def test_philox_mode_equals_replay_of_its_own_draws_merton_random_parameters(cuda_lib, sampler):
    """The same with every model parameter drawn at random - in particular T != 1, on which the step search,
    the block index of a jump and the end-of-path test of the three walks depend."""
    rng = np.random.default_rng(2718)
    classes = (mm.Euro_Merton, mm.Asian_Merton, mm.Lookback_Merton, mm.Digital_Merton)
    levels = ((0, 2), (1, 2), (2, 2), (2, 4), (3, 4), (2, 5), (7, 2), (4, 4))
    for it in range(16):
        kw = dict(X0=float(rng.uniform(20, 300)), T=float(rng.uniform(0.2, 3.0)), r=float(rng.uniform(0.0, 0.1)),
                  sig=float(rng.uniform(0.05, 0.6)), lam=float(rng.uniform(0.2, 8.0)),
                  jumpmean=float(rng.uniform(-0.3, 0.3)), jumpstd=float(rng.uniform(0.05, 0.5)))
        kw["K"] = kw["X0"] * float(rng.uniform(0.7, 1.3))
        cls = classes[it % 4]
        l, M = levels[it % len(levels)]
        _merton_own_draws_case(cuda_lib, cls, kw, l, M, n=300, off=int(rng.integers(0, 2 ** 40)), seed=int(rng.integers(0, 2 ** 31)),
                               K=128, sampler=sampler)


def test_merton_walks_agree_per_path(cuda_lib, tmp_path):
    """The block-uniform (ring of records per thread), the event-driven (lane queues) and the batch (records of a
    whole warp dealt out one per lane) Merton walks consume the same draws through the same arithmetic: per-path
    payoffs are identical bit for bit, whatever the length of the event-driven walk's advance phase.  The factor
    walk (the default for Euler-Maruyama, M = 2 | 4) consumes the same draws with the jump arithmetic in factor
    form: the same payoffs to rounding (1e-12 relative), including the cases that need several record rounds
    (lam*T = 12 and 40)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("factor", {}),
                     ("uniform", {"MLMCB_MJD_FLAT_BELOW": "0", "MLMCB_MJD_WALK": "uniform"}),
                     ("event", {"MLMCB_MJD_FLAT_BELOW": "1000000", "MLMCB_MJD_WALK": "event"}),
                     ("batch", {"MLMCB_MJD_FLAT_BELOW": "1000000", "MLMCB_MJD_WALK": "batch"}),
                     ("event_run1", {"MLMCB_MJD_FLAT_BELOW": "1000000", "MLMCB_MJD_WALK": "event", "MLMCB_MJD_RUN_BLOCKS": "1"}),
                     ("event_run3", {"MLMCB_MJD_FLAT_BELOW": "1000000", "MLMCB_MJD_WALK": "event", "MLMCB_MJD_RUN_BLOCKS": "3"})):
        f = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, os.path.join(root, "tests", "merton_walks_case.py"), f], check=True, cwd=root,
                       env={**os.environ, **env, "PYTHONPATH": root})
        res[tag] = dict(np.load(f))
    for tag in ("event", "batch", "event_run1", "event_run3"):
        for k, v in res["uniform"].items():
            assert np.array_equal(v, res[tag][k]), (tag, k, np.abs(v - res[tag][k]).max())
    differ = 0
    for k, v in res["uniform"].items():
        np.testing.assert_allclose(res["factor"][k], v, rtol=1e-12, atol=1e-10, err_msg=k)
        differ += int(not np.array_equal(res["factor"][k], v))
    assert differ > 0       # the factor walk did run (its rounding is its own)


def test_merton_without_jumps_is_gbm(cuda_lib):
    """lam = 0: no record ever falls inside the path (infinite arrival times), the drift is r, and the grid
    normals are addressed like the GBM kernel's - so every Merton walk reproduces the GBM payoffs."""
    for mer, gbm in ((mm.Euro_Merton, mm.Euro_GBM), (mm.Asian_Merton, mm.Asian_GBM), (mm.Lookback_Merton, mm.Lookback_GBM)):
        for l, M in ((2, 2), (3, 4), (4, 4), (2, 3)):       # 4, 64 steps event-driven; 256 block-uniform; 9 generic M
            a, b = mer(lam=0.0, sig=0.3, seed=5), gbm(sig=0.3, seed=5)
            a._next_path[l] = b._next_path[l] = 1234
            Pf, Pc = a.payoff(3001, l, M)
            Gf, Gc = b.payoff(3001, l, M)
            np.testing.assert_allclose(Pf, Gf, rtol=1e-13, atol=1e-12)
            np.testing.assert_allclose(Pc, Gc, rtol=1e-13, atol=1e-12)
        s = mer(lam=0.0, sig=0.3, seed=5).looper(100_000, 0, 2, False)      # l = 0: own draw mapping, finite sums
        assert np.isfinite(s).all() and s[0] == s[2]


def test_jump_stream_distributions(cuda_lib):
    """Jump records: inter-arrival times ~ Exp(lam) (:1294), jump-size and sub-step normals ~ N(0,1),
    mutually uncorrelated, and independent of the grid-step normals of the same path."""
    import torch
    from scipy import stats
    lam, n, K = 2.5, 1 << 17, 8
    for sampler in (_lib.SAMPLER_F32BM, _lib.SAMPLER_F64BM):
        rec = torch.empty(3 * K * n, dtype=torch.float64, device="cuda")
        _lib.check(cuda_lib.mlmcb_jump_probe(sampler, 7, 2, lam, 0, n, K, 0, 0, rec.data_ptr()))
        torch.cuda.synchronize()
        gap, zq, zw = rec.cpu().numpy().reshape(3, K * n)
        N = gap.size
        assert abs(gap.mean() * lam - 1) < 4 / np.sqrt(N)
        assert abs(gap.var() * lam ** 2 - 1) < 4 * np.sqrt(8 / N)
        assert stats.kstest(gap[:400000] * lam, "expon").pvalue > 1e-3
        assert gap.min() > 0
        for z in (zq, zw):
            assert abs(z.mean()) < 4 / np.sqrt(N) and abs(z.var() - 1) < 4 * np.sqrt(2 / N)
        for a, b in ((gap, zq), (gap, zw), (zq, zw)):
            assert abs(np.corrcoef(a, b)[0, 1]) < 4 / np.sqrt(N)
        grid = _probe_normals(cuda_lib, sampler, 7, 2, 0, n, K).reshape(-1)
        assert abs(np.corrcoef(grid, zw)[0, 1]) < 4 / np.sqrt(N)


@pytest.mark.parametrize("sampler", ["f32bm", "f64bm"])
def test_packed_levels_equal_replay_of_their_own_draws(cuda_lib, sampler):
    """l = 0 (hand-specialised loop, 4 ids per Philox block) and M=2, l=1 (2 ids per block): the normals
    from the packed probe replayed through the parity-checked path give the same per-path payoffs."""
    import torch
    code = _lib.SAMPLER_F32BM if sampler == "f32bm" else _lib.SAMPLER_F64BM
    for cls in (mm.Euro_GBM, mm.Asian_GBM, mm.Lookback_GBM, mm.Digital_GBM):
        for l, M in ((0, 2), (0, 4), (0, 3), (1, 2)):
            n, off, seed = 1003, 5, 9
            nsteps = M ** l
            opt = cls(seed=seed, sampler=sampler)
            opt._next_path[l] = off
            Pf, Pc = opt.payoff(n, l, M)
            out = torch.empty(nsteps * n, dtype=torch.float64, device="cuda")
            _lib.check(cuda_lib.mlmcb_packed_probe(code, seed, l, nsteps, off, n, 0, 0, out.data_ptr()))
            torch.cuda.synchronize()
            Z = out.cpu().numpy().reshape(nsteps, n)
            Rf, Rc = cls().payoff_replay(Z, l, M, arith="fast")
            np.testing.assert_allclose(Pf, Rf, rtol=1e-13, atol=1e-11)
            np.testing.assert_allclose(Pc, Rc, rtol=1e-13, atol=1e-11)


def test_torch_custom_op(cuda_lib):
    import torch
    from multilevel_mc_sdes_b200 import torch_ops
    opt = mm.Lookback_GBM(seed=3)
    want = mm.Lookback_GBM(seed=3).looper(100_000, 3, 4, False)
    got = torch.ops.mlmcb.level_sums(torch_ops.option_params(opt), opt._model, opt._payoff_kind, 3, 4, 100_000, 3, 0, 0, 0)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("sampler", ["f64bm", "f32bm"])
def test_sweep_kernel_equals_single_level_calls(cuda_lib, sampler):
    """One launch for all levels (sweep_kernel: per-level block ranges, bodies of the packed / four-step / long GBM
    kernels and of the three Merton walks) against one level-kernel launch per level on the same path ids: equal up
    to fp64 summation order - every product, both schemes, M = 2, 4 and a generic M, ragged sample counts, an empty
    request in the middle, and a second sweep on the same object (fresh ids, workspace reuse)."""
    for cname in sorted(CLASS_CODES):
        for M, Lmax, scheme in ((2, 8, "em"), (4, 4, "em"), (3, 3, "em"), (4, 3, "milstein")):
            a = getattr(mm, cname)(seed=12, sampler=sampler, scheme=scheme, verbose=False)
            b = getattr(mm, cname)(seed=12, sampler=sampler, scheme=scheme, verbose=False)
            for rep in range(2):
                reqs = [(1000 + 37 * l + 5000 * (l == 0) + rep, l) for l in range(Lmax + 1)]
                reqs[2] = (0, 2)                                   # nothing asked of level 2
                rows = a._sweep(reqs, M)
                for (n, l), row in zip(reqs, rows):
                    if n == 0:
                        assert not row.any()
                        b._take_ids(l, 0)
                        continue
                    want = b.looper(n, l, M, False)
                    np.testing.assert_allclose(row, want, rtol=2e-12, atol=1e-9, err_msg=f"{cname} M={M} l={l} {scheme}")
