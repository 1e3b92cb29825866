"""CPU: host logic of the drop-in surface, the C-ABI export list, Philox known answers."""
import ctypes
import os
import re

import numpy as np
import pytest

import multilevel_mc_sdes_b200 as mm
from multilevel_mc_sdes_b200 import _lib
from tests.conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "mlmcb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlmcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mlmcb.h but not exported"
    assert sorted(_lib.SIGNATURES) == names       # the Python binding covers the whole header


def test_params_struct_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "mlmcb.h")).read()
    body = re.search(r"typedef struct mlmcb_params \{(.*?)\} mlmcb_params;", text, re.S).group(1)
    fields = re.findall(r"\b([A-Za-z_0-9]+)\s*[,;]", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f for f, _ in _lib.Params._fields_]
    # 13 doubles, one pointer, two ints
    assert ctypes.sizeof(_lib.Params) == 13 * 8 + 8 + 2 * 4 and len(fields) == 16


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert _lib.philox4x32_10(ctr, key) == out


def test_fill_derived():
    p = _lib.Params(X0=100, K=100, T=2.0, r=0.05, sig=0.2, jumpmean=0.1, jumpstd=0.2)
    _lib.load().mlmcb_params_fill_derived(ctypes.byref(p))
    assert p.disc == pytest.approx(np.exp(-0.1), rel=1e-15)
    assert p.J_bar == pytest.approx(np.exp(0.1 + 0.02) - 1, rel=1e-15)


def test_closed_forms_match_reference(golden):
    g = golden["closed_forms"]
    assert mm.Euro_GBM().BS() == pytest.approx(float(g["Euro_GBM"]), rel=1e-14)
    assert mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5).BS() == pytest.approx(35.17419437739385, rel=1e-14)
    assert mm.Asian_GBM().BS() == pytest.approx(float(g["Asian_GBM"]), rel=1e-14)
    assert mm.Lookback_GBM().BS() == pytest.approx(float(g["Lookback_GBM"]), rel=1e-14)
    assert mm.Digital_GBM().BS() == pytest.approx(float(g["Digital_GBM"]), rel=1e-14)
    assert mm.Euro_Merton().BS() == pytest.approx(float(g["Euro_Merton"]), rel=1e-13)
    assert mm.Euro_Merton(lam=3, jumpmean=-0.05, jumpstd=0.3, sig=0.3, X0=110).BS(n=60) == pytest.approx(
        float(g["Euro_Merton_alt"]), rel=1e-13)
    assert mm.Euro_GBM(X0=90, K=100, r=0.03, sig=0.35, T=2.0, drift=0.01).BS() == pytest.approx(
        float(g["Euro_GBM_alt2"]), rel=1e-14)
    with pytest.raises(NotImplementedError):
        mm.Asian_Merton().BS()


def test_mlmc_driver_replays_reference_traces(golden):
    """The host driver, fed the 7-vectors the reference's looper returned, must ask for the same
    (Nl, l) sequence and return the same (sums, N) bit for bit (MLMC_RSCAM.py:960-1042)."""
    t = golden["mlmc_trace"]
    for i in range(int(t["n_cases"])):
        meta = [str(x) for x in t[f"t{i}_meta"]]       # the anti=True trace drives the same host loop
        kw = eval(meta[1])
        eps, M, N0 = float(meta[2]), int(meta[3]), int(meta[4])
        req, ret = t[f"t{i}_req"], t[f"t{i}_ret"]
        pos = [0]

        def sweep(requests, M_):
            out = []
            for n, l in requests:
                assert (n, l) == tuple(req[pos[0]])
                out.append(ret[pos[0]])
                pos[0] += 1
            return out

        sums, N, a, b = mm.mlmc_driver(sweep, eps, M, N0, kw.get("alpha_0"), kw.get("beta_0"))
        assert pos[0] == len(req)
        assert np.array_equal(sums, t[f"t{i}_sums"]) and np.array_equal(N, t[f"t{i}_N"])
        assert [a, b] == list(t[f"t{i}_ab"])
        assert N.dtype == np.float64 and sums.shape == (7, len(N))


def synthetic_sweep(alpha=1.0, beta=2.0, c1=1.0, c2=1.0):
    """Deterministic stand-in level function with E[dP_l] ~ M^-alpha l, V ~ M^-beta l."""
    def sweep(requests, M):
        out = []
        for n, l in requests:
            m, v = c1 * M ** (-alpha * l), c2 * M ** (-beta * l)
            out.append(np.array([n * m, n * (v + m * m), n * 10.0, n * 101.0, 0, 0, 0]))
        return out
    return sweep


def test_mlmc_none_safe_and_regression_branch():
    # alpha_0 = beta_0 = None (the README call) must run: regress alpha, beta each iteration
    sums, N, a, b = mm.mlmc_driver(synthetic_sweep(1.0, 2.0), 0.01, 2, 1000, None, None)
    assert a == pytest.approx(1.0, abs=1e-9) and b == pytest.approx(2.0, abs=1e-9)
    assert len(N) >= 3 and np.all(N >= 1000)
    # optimal allocation: N_l proportional to sqrt(V_l / C_l)
    assert np.all(np.diff(N) <= 0) and N[0] > 10 * N[-1]


def test_mlmc_floor_half_and_zero_variance_fix():
    # variance identically zero above level 2 (Digital-like): V[3:] extrapolated from V[2], no NaN
    def sweep(requests, M):
        out = []
        for n, l in requests:
            v = 1.0 if l <= 2 else 0.0
            m = 0.5 * M ** (-1.0 * l)
            out.append(np.array([n * m, n * (v + m * m), 0, 0, 0, 0, 0]))
        return out
    sums, N, a, b = mm.mlmc_driver(sweep, 0.02, 2, 1000, 1, None)
    assert np.all(np.isfinite(N)) and b >= 0.5


def test_option_surface_and_errors():
    o = mm.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5)
    assert (o.X0, o.K, o.r, o.sig, o.T, o.alpha_0, o.beta_0) == (125, 105, 0.02, 0.5, 1, None, None)
    assert o.mu == 0.02
    m = mm.Digital_Merton()
    assert m.lam == 1 and m.jumpmean == 0.1 and m.jumpstd == 0.2
    assert m.J_bar == np.exp(0.1 + 0.5 * 0.2 ** 2) - 1
    assert mm.Lookback_GBM.beta == 0.5826 and mm.Lookback_Merton.beta == 0.5826
    with pytest.raises(NotImplementedError):
        mm.Lookback_GBM().looper(10, 1, 2, True)            # anti on a product without anti_payoff (:844)
    with pytest.raises(NotImplementedError):
        mm.Euro_Merton().mlmc(0.1, anti=True)
    with pytest.raises(NotImplementedError):
        mm.Digital_GBM().anti_payoff(10, 1, 2)
    assert mm.Euro_GBM._has_anti and mm.Asian_GBM._has_anti and not mm.Lookback_GBM._has_anti
    with pytest.raises(NotImplementedError):
        mm.Option().looper(10, 1, 2, False)                  # base class has no payoff (:829)
    with pytest.raises(ValueError):
        mm.Euro_GBM(sampler="ziggurat")
    x = np.array([100.0, 101.0])
    assert np.array_equal(o.sde(x, 0.01, 0.0, np.array([0.1, -0.1])), o.mu * x * 0.01 + o.sig * x * np.array([0.1, -0.1]))


def test_reference_class_names_are_all_present():
    """Every class the reference module defines exists here (MLMC_RSCAM.py:13-16); the user-lambda
    ones refuse to construct instead of silently running something else."""
    for name in ("Option", "JumpDiffusion_Option", "Diffusion_Option", "Merton_Option", "GBM_Option", "MyOption",
                 "Euro_GBM", "Euro_Merton", "Lookback_GBM", "Lookback_Merton", "Asian_GBM", "Asian_Merton",
                 "Digital_GBM", "Digital_Merton", "Amer_GBM"):
        assert hasattr(mm, name), name
    for cls in (mm.JumpDiffusion_Option, mm.Diffusion_Option, mm.MyOption):
        with pytest.raises(NotImplementedError):
            cls(sig=lambda x, t: 0.2)
    m = mm.Euro_Merton()
    assert callable(m.jumptime) and callable(m.jumpsize)


def test_amer_gbm_surface():
    a = mm.Amer_GBM(exerciseBoundary=mm.sqrt_boundary(0.7), K=95)
    assert a.boundary_c == 0.7 and a._params().boundary_c == 0.7
    assert a.exerciseBoundary(a, 0.5) == max(95 * (1 - 0.7 * np.sqrt(0.5)), 0)
    assert mm.Amer_GBM().boundary_c == 0.8 and mm.Amer_GBM(0.6).boundary_c == 0.6
    assert a._params((3,)).amer_levels == 0                      # the family is tabulated on the device
    # any other callable (the reference's contract, :1495-1497) is evaluated on the host, once per level, at the
    # reference's times: grid points, fine mid-points, coarse mid-points
    calls = []
    g = mm.Amer_GBM(exerciseBoundary=lambda self, t: calls.append(t) or self.K * (0.8 + 0.1 * t), K=90, T=2.0)
    p = g._params((0, 2))
    assert p.amer_levels == 3 and not p.amer_boundary[1] and p.amer_boundary[0] and p.amer_boundary[2]
    tab = g._boundary_values(2)
    dt = 2.0 / 4
    want_t = [j * dt for j in range(5)] + [(j - 1) * dt + dt / 2 for j in range(1, 5)] + [(j - 2) * dt + 2 * dt / 2 for j in (2, 4)]
    assert np.array_equal(tab, [90 * (0.8 + 0.1 * t) for t in want_t])
    n_calls = len(calls)
    g._params((2,))
    assert len(calls) == n_calls                                 # cached per level


def test_path_ids_advance_per_level(monkeypatch):
    """Every looper call must consume fresh path ids."""
    seen, flags_seen = [], []

    class FakeLib:
        def mlmcb_level_sweep_host(self, p, model, payoff, M, k, levels, n, off, seed, flags, dev, st, out):
            seen.append([(levels[i], n[i], off[i]) for i in range(k)])
            flags_seen.append(flags)
            return 0

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())
    o = mm.Asian_GBM(seed=7)
    o.looper(1000, 3, 4, False)
    o.looper(500, 3, 4, False)
    o.looper(200, 2, 4, False)
    o.level_sweep(Lmax=3, Nsamples=10, M=4)
    assert flags_seen[:3] == [0, 0, 0]
    o.looper(5, 3, 4, True)
    assert flags_seen[-1] == _lib.ANTITHETIC and seen[-1] == [(3, 5, 1510)]
    assert seen[0] == [(3, 1000, 0)] and seen[1] == [(3, 500, 1000)] and seen[2] == [(2, 200, 0)]
    assert seen[3] == [(0, 10, 0), (1, 10, 0), (2, 10, 200), (3, 10, 1500)]


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.MlmcbError):
        mm.Euro_GBM().looper(100, 1, 2, False)


def test_invalid_arguments_are_value_errors():
    import torch
    if torch.cuda.is_available():
        with pytest.raises(ValueError):
            mm.Euro_GBM().looper(100, 1, 1, False)           # M < 2


def test_flag_combinations_are_checked_before_the_device():
    """Flag validation needs no GPU: unknown bits are EINVAL, combinations that were never built are ENOTIMPL."""
    import ctypes as C
    lib = _lib.load()
    t, b = C.c_int(), C.c_int()
    geo = lambda flags: lib.mlmcb_launch_geometry(0, 0, 4, flags, 0, C.byref(t), C.byref(b))
    assert geo(0x1000) == _lib.EINVAL
    assert geo(0xE) == _lib.EINVAL                                              # unknown sampler code
    assert geo(_lib.FP32_PATHS | _lib.MILSTEIN) == _lib.ENOTIMPL               # Milstein: fp64 path arithmetic only
    assert b"Milstein" in lib.mlmcb_last_error()
    assert _lib.SAMPLER_F64BM == 0                                              # flags == 0 is the all-fp64 kernel


def test_f64_sampler_tables_match_their_generator():
    """csrc/mlmcb_f64tab.inc and the polynomial coefficients compiled into mlmcb_api.cu are exactly what
    tools/gen_f64bm_tables.py produces (60-digit arithmetic, rounded once), and a sample of pairs emulated in
    float64 / long-double FMA stays within 2e-15 of the 60-digit normals."""
    pytest.importorskip("mpmath")
    import re
    import sys
    tools = os.path.join(ROOT, "tools")
    sys.path.insert(0, tools)
    try:
        import gen_f64bm_tables as g
    finally:
        sys.path.remove(tools)
    lg, tr = g.log_table(), g.trig_table()
    lq, ts, tc = g.coefficients()
    inc = open(os.path.join(ROOT, "multilevel_mc_sdes_b200", "csrc", "mlmcb_f64tab.inc")).read()
    rows = re.findall(r"\{(-?0x[0-9a-f.]+p[+-]\d+), (-?0x[0-9a-f.]+p[+-]\d+)\}", inc)
    assert len(rows) == len(lg) + len(tr) == 512
    want = [(a.hex(), b.hex()) for a, b in lg + tr]
    assert rows == want
    api = open(os.path.join(ROOT, "multilevel_mc_sdes_b200", "csrc", "mlmcb_api.cu")).read()
    for c in lq + ts + tc:
        assert c.hex() in api, c.hex()
    assert g.check(lg, tr, lq, ts, tc, n=1500) < 2e-15
