"""GPU: replay parity - the CUDA path fed the reference's own draws (north-star check (a)).

Through the C ABI (mlmcb_replay_gbm / mlmcb_replay_merton) on every golden case generated from the
unmodified reference.  Bar: GBM with the reference-order arithmetic policy is BIT-EXACT; the
production (FMA) policy and Merton are within 1e-12 * max(|P|, K e^{-rT})."""
import numpy as np
import pytest

import multilevel_mc_sdes_b200 as mm
from oracle import np_oracle as O
from tests.helpers import amer_cases, milstein_gbm_cases, milstein_merton_cases
from tests.helpers import CLASS_CODES, anti_cases, case_kwargs, gbm_cases, merton_cases, oracle_params, payoff_tol, regen_Z

pytestmark = pytest.mark.gpu


def _opt(cname, kw_name):
    return getattr(mm, cname)(**case_kwargs(cname, kw_name))


def test_gbm_replay_exact_policy_is_bit_exact(golden, cuda_lib):
    n = 0
    for c in gbm_cases(golden["payoffs_gbm"]):
        Pf, Pc = _opt(c["cname"], c["kw_name"]).payoff_replay(regen_Z(c), c["l"], c["M"], arith="exact")
        assert np.array_equal(Pf, c["Pf"]), (c["cname"], c["kw_name"], c["l"], c["M"])
        assert np.array_equal(Pc, c["Pc"]), (c["cname"], c["kw_name"], c["l"], c["M"])
        n += 1
    assert n > 100


def test_gbm_replay_production_policy_within_1e12(golden, cuda_lib):
    worst = 0.0
    for c in gbm_cases(golden["payoffs_gbm"]):
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = _opt(c["cname"], c["kw_name"]).payoff_replay(regen_Z(c), c["l"], c["M"], arith="fast")
        if c["cname"] == "Digital_GBM":
            # an indicator may flip only if X is within rounding of K (astronomically rare)
            assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"])
            continue
        ef, ec = np.abs(Pf - c["Pf"]), np.abs(Pc - c["Pc"])
        assert np.all(ef <= payoff_tol(p, c["Pf"])), (c["cname"], c["l"], c["M"], ef.max())
        assert np.all(ec <= payoff_tol(p, c["Pc"])), (c["cname"], c["l"], c["M"], ec.max())
        worst = max(worst, (ef / payoff_tol(p, c["Pf"])).max())
    assert worst < 1.0


@pytest.mark.parametrize("arith", ["exact", "fast"])
def test_merton_replay_within_1e12(golden, cuda_lib, arith):
    n = 0
    for c in merton_cases(golden["payoffs_merton"]):
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = _opt(c["cname"], c["kw_name"]).payoff_replay(c["packed"], c["l"], c["M"], arith=arith)
        if c["cname"] == "Digital_Merton":
            assert np.array_equal(Pf, c["Pf"])
            if c["l"] == 0:      # second item is the raw coarse value (:133), which jumps do move
                assert np.all(np.abs(Pc - c["Pc"]) <= 1e-12 * np.abs(c["Pc"]))
            else:
                assert np.array_equal(Pc, c["Pc"])
        else:
            assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])), (c["idx"], np.abs(Pf - c["Pf"]).max())
            assert np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"])), (c["idx"], np.abs(Pc - c["Pc"]).max())
        n += 1
    assert n > 60


def test_replay_large_level_against_oracle(cuda_lib, c_oracle):
    """l=8, M=4 (65 536 fine steps) and l=10, M=2: sizes past the fixtures, checked against the
    pinned C restatement on the same seeded normals."""
    for (l, M, n) in ((8, 4, 96), (10, 2, 128)):
        rs = np.random.RandomState(1234 + l)
        Z = rs.randn(M ** l, n)
        for cname, pay in (("Euro_GBM", O.EURO), ("Asian_GBM", O.ASIAN), ("Lookback_GBM", O.LOOKBACK)):
            p = O.Params()
            ref_f, ref_c = c_oracle.gbm_replay(p, pay, l, M, Z)
            opt = getattr(mm, cname)()
            Pf, Pc = opt.payoff_replay(Z, l, M, arith="exact")
            assert np.array_equal(Pf, ref_f) and np.array_equal(Pc, ref_c)
            Pf, Pc = opt.payoff_replay(Z, l, M, arith="fast")
            assert np.all(np.abs(Pf - ref_f) <= payoff_tol(p, ref_f))
            assert np.all(np.abs(Pc - ref_c) <= payoff_tol(p, ref_c))


def test_replay_sums_row_order(cuda_lib, golden):
    """d_out7 of the replay entry point follows looper's row order (:955) and l==0 convention (:949)."""
    import ctypes as C
    import torch
    from multilevel_mc_sdes_b200 import _lib
    for c in list(gbm_cases(golden["payoffs_gbm"]))[:12]:
        opt = _opt(c["cname"], c["kw_name"])
        Z = torch.as_tensor(regen_Z(c), device="cuda")
        n = c["n"]
        Pf = torch.empty(n, dtype=torch.float64, device="cuda")
        Pc = torch.empty(n, dtype=torch.float64, device="cuda")
        out = torch.zeros(7, dtype=torch.float64, device="cuda")
        p = opt._params()
        _lib.check(cuda_lib.mlmcb_replay_gbm(C.byref(p), opt._payoff_kind, c["l"], c["M"], n, Z.data_ptr(),
                                             _lib.ARITH_EXACT, 0, 0, Pf.data_ptr(), Pc.data_ptr(), out.data_ptr()))
        torch.cuda.synchronize()
        want = O.seven_sums(c["Pf"], c["Pc"], c["l"])
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-13, atol=1e-9)


def test_replay_edge_cases(cuda_lib):
    opt = mm.Euro_GBM()
    # a single path, l=0
    Pf, Pc = opt.payoff_replay(np.array([[0.3]]), 0, 2, arith="exact")
    p = O.Params()
    ref = O.level_payoffs(p, O.GBM, O.EURO, 0, 2, Z=np.array([[0.3]]))
    assert Pf[0] == ref[0][0] and Pc[0] == 100.0
    # wrong number of rows is rejected before any launch
    with pytest.raises(ValueError):
        opt.payoff_replay(np.zeros((3, 4)), 2, 2)
    # Merton path with no jumps at all before T and one with several
    m = mm.Euro_Merton()
    packed = dict(zW=np.array([0.1, -0.2, 0.3, 0.4, 0.5, 0.0, -1.0, 0.7]), offW=np.array([0, 2, 8], dtype=np.int64),
                  zQ=np.array([0.2, -0.3, 0.1, 1.0]), offQ=np.array([0, 0, 4], dtype=np.int64),
                  E=np.array([5.0, 0.1, 0.3, 0.2, 0.05, 9.0]), offE=np.array([0, 1, 6], dtype=np.int64))
    Pf, Pc = m.payoff_replay(packed, 1, 2, arith="exact")
    want = O.level_payoffs(O.Params(), O.MERTON, O.EURO, 1, 2, n=2, draws=O.ReplayDraws(packed))
    np.testing.assert_allclose(Pf, want[0], rtol=1e-13)
    np.testing.assert_allclose(Pc, want[1], rtol=1e-13)


def test_antithetic_replay(golden, cuda_lib, c_oracle):
    """anti_payoff (:292-317): reference-order policy bit-exact, production policy within 1e-12."""
    n = 0
    for c in anti_cases(golden["payoffs_anti"]):
        p = oracle_params(c["cname"], c["kw_name"])
        opt = _opt(c["cname"], c["kw_name"])
        Z = regen_Z(c)
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="exact", anti=True)
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), (c["cname"], c["l"], c["M"])
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="fast", anti=True)
        assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])) and np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"]))
        n += 1
    assert n > 40
    # beyond the fixtures: l=7, M=4 (16 384 steps) against the pinned C restatement
    Z = np.random.RandomState(77).randn(4 ** 7, 64)
    for cname, pay in (("Euro_GBM", O.EURO), ("Asian_GBM", O.ASIAN)):
        rf, rc = c_oracle.gbm_anti_replay(O.Params(), pay, 7, 4, Z)
        Pf, Pc = getattr(mm, cname)().payoff_replay(Z, 7, 4, arith="exact", anti=True)
        assert np.array_equal(Pf, rf) and np.array_equal(Pc, rc)
    with pytest.raises(ValueError, match="odd coarseness factor"):
        mm.Euro_GBM().payoff_replay(np.zeros((9, 4)), 2, 3, anti=True)
    with pytest.raises(NotImplementedError):
        mm.Lookback_GBM().payoff_replay(np.zeros((4, 4)), 2, 2, anti=True)


def test_antithetic_replay_with_milstein_stepper(golden, cuda_lib):
    """scheme="milstein" + anti=True (the reference composes them: diffusion_anti_path calls whatever self.sde is
    bound to, :200-210; the demo notebook binds milstein_GBM): reference-order policy bit-exact, production 1e-12."""
    n = 0
    for c in anti_cases(golden["payoffs_anti_milstein"]):
        p = oracle_params(c["cname"], c["kw_name"])
        opt = getattr(mm, c["cname"])(scheme="milstein", **case_kwargs(c["cname"], c["kw_name"]))
        Z = regen_Z(c)
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="exact", anti=True)
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), (c["cname"], c["l"], c["M"])
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="fast", anti=True)
        assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])) and np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"]))
        n += 1
    assert n > 40


def test_merton_replay_deep_levels(golden, cuda_lib):
    """Reference fixtures at 128 and 256 fine steps (l = 7, 8 at M = 2; l = 4 at M = 4) and at 512 and 1024 (l = 9 at
    M = 2; l = 5 at M = 4), both policies."""
    n = 0
    for c in merton_cases(golden["payoffs_merton_deep"]):
        p = oracle_params(c["cname"], c["kw_name"])
        for arith in ("exact", "fast"):
            Pf, Pc = _opt(c["cname"], c["kw_name"]).payoff_replay(c["packed"], c["l"], c["M"], arith=arith)
            if c["cname"] == "Digital_Merton":
                assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"])
            else:
                assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])), (c["idx"], arith, np.abs(Pf - c["Pf"]).max())
                assert np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"])), (c["idx"], arith, np.abs(Pc - c["Pc"]).max())
        n += 1
    assert n == 28


def test_milstein_replay(golden, cuda_lib):
    """scheme="milstein": GBM reference-order policy bit-exact, everything else within 1e-12."""
    g = golden["payoffs_milstein"]
    n = 0
    for c in milstein_gbm_cases(g):
        p = oracle_params(c["cname"], c["kw_name"])
        opt = getattr(mm, c["cname"])(scheme="milstein", **case_kwargs(c["cname"], c["kw_name"]))
        Z = regen_Z(c)
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="exact")
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), (c["cname"], c["l"], c["M"])
        Pf, Pc = opt.payoff_replay(Z, c["l"], c["M"], arith="fast")
        if c["cname"] == "Digital_GBM":
            assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"])
        else:
            assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])) and np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"]))
        n += 1
    for c in milstein_merton_cases(g):
        p = oracle_params(c["cname"], c["kw_name"])
        opt = getattr(mm, c["cname"])(scheme="milstein", **case_kwargs(c["cname"], c["kw_name"]))
        for arith in ("exact", "fast"):
            Pf, Pc = opt.payoff_replay(c["packed"], c["l"], c["M"], arith=arith)
            if c["cname"] == "Digital_Merton":
                assert np.array_equal(Pf, c["Pf"])
            else:
                assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])), (c["idx"], arith)
                assert np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"])), (c["idx"], arith)
        n += 1
    assert n > 100


def test_american_put_replay(golden, cuda_lib, c_oracle):
    """Amer_GBM: both arithmetic policies within 1e-12*max(|P|, K e^{-rT}) of the reference (the
    per-step exp() makes bit-exactness impossible across libms); l = 9 against the C restatement."""
    n = 0
    for c in amer_cases(golden["payoffs_amer"]):
        kw = case_kwargs("Euro_GBM", c["kw_name"])
        p = O.Params(boundary_c=c["c"], **kw)
        opt = mm.Amer_GBM(exerciseBoundary=c["c"], **kw)
        Z = regen_Z(c)
        for arith in ("exact", "fast"):
            Pf, Pc = opt.payoff_replay(Z, c["l"], 2, arith=arith)
            assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])), (c["idx"], arith, np.abs(Pf - c["Pf"]).max())
            assert np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"])), (c["idx"], arith, np.abs(Pc - c["Pc"]).max())
        n += 1
    assert n >= 24
    Z = np.random.RandomState(5).randn(2 ** 9, 80)
    rf, rc = c_oracle.gbm_replay(O.Params(), O.AMERICAN, 9, 2, Z)
    Pf, Pc = mm.Amer_GBM().payoff_replay(Z, 9, 2, arith="exact")
    assert np.all(np.abs(Pf - rf) <= payoff_tol(O.Params(), rf)) and np.all(np.abs(Pc - rc) <= payoff_tol(O.Params(), rc))
    with pytest.raises(ValueError, match="M must be equal to 2"):
        mm.Amer_GBM().payoff_replay(np.zeros((16, 4)), 2, 4)


def test_american_put_with_tabulated_boundary_matches_reference(cuda_lib, golden):
    """Amer_GBM(exerciseBoundary=<any callable>), the reference's contract (:1495-1497): the callable is evaluated
    on the host at the reference's times and only the values enter the library.  Fixtures: the reference run with
    the two boundaries of oracle/boundaries.py (neither is of the square-root family)."""
    from oracle.boundaries import BOUNDARIES
    n = 0
    g = golden["payoffs_amer_callable"]
    for i in range(int(g["n_cases"])):
        cname, kw_name, l, M, npaths, seed, bname = [str(x) for x in g[f"c{i}_meta"]]
        c = dict(l=int(l), M=int(M), n=int(npaths), seed=int(seed), Z=None)
        kw = case_kwargs("Euro_GBM", kw_name)
        p = O.Params(**kw)
        opt = mm.Amer_GBM(exerciseBoundary=BOUNDARIES[bname], **kw)
        Z = regen_Z(c)
        for arith in ("exact", "fast"):
            Pf, Pc = opt.payoff_replay(Z, c["l"], 2, arith=arith)
            assert np.all(np.abs(Pf - g[f"c{i}_Pf"]) <= payoff_tol(p, g[f"c{i}_Pf"])), (i, arith, np.abs(Pf - g[f"c{i}_Pf"]).max())
            assert np.all(np.abs(Pc - g[f"c{i}_Pc"]) <= payoff_tol(p, g[f"c{i}_Pc"])), (i, arith, np.abs(Pc - g[f"c{i}_Pc"]).max())
        n += 1
    assert n == 20
    # the tabulated path and the device-tabulated family agree when the callable IS the family
    fam = lambda self, t: np.maximum(self.K * (1 - 0.8 * np.sqrt(self.T - t)), 0)       # the notebook's exb, unwrapped
    Z = np.random.RandomState(3).randn(2 ** 5, 200)
    a = mm.Amer_GBM(exerciseBoundary=fam).payoff_replay(Z, 5, 2, arith="exact")
    b = mm.Amer_GBM(exerciseBoundary=0.8).payoff_replay(Z, 5, 2, arith="exact")
    for x, y in zip(a, b):
        assert np.all(np.abs(x - y) <= payoff_tol(O.Params(), y))
    # production kernels: sums, per-path payoffs, paths and a sweep all take the table
    opt = mm.Amer_GBM(exerciseBoundary=BOUNDARIES["power"], seed=5)
    s = opt.looper(200_000, 4, 2, False)
    opt2 = mm.Amer_GBM(exerciseBoundary=BOUNDARIES["power"], seed=5)
    Pf, Pc = opt2.payoff(200_000, 4, 2)
    np.testing.assert_allclose(s, O.seven_sums(Pf, Pc, 4), rtol=1e-12)
    sw = mm.Amer_GBM(exerciseBoundary=BOUNDARIES["linear"], seed=6).level_sweep(Lmax=3, Nsamples=20_000, M=2)
    assert sw.shape == (7, 4) and np.isfinite(sw).all() and sw[2, 0] > 0
    F, Cc = mm.Amer_GBM(exerciseBoundary=BOUNDARIES["linear"], seed=6).path(1000, 3, 2)
    assert F.shape == (1000, 3) and Cc.shape == (1000, 3)


def test_randomised_parameters_bit_exact_against_c_oracle(cuda_lib, c_oracle):
    """40 seeded random parameter sets / shapes (odd path counts, M in 2..6, every GBM product, plain,
    Milstein and antithetic): the reference-order policy must equal the pinned C restatement bit for bit,
    the production policy within 1e-12*max(|P|, K e^{-rT})."""
    rs = np.random.RandomState(20261020)
    n_checked = 0
    for case in range(40):
        kw = dict(X0=float(rs.uniform(50, 150)), K=float(rs.uniform(60, 140)), T=float(rs.choice([0.25, 0.5, 1.0, 2.0, 3.0])),
                  r=float(rs.uniform(0.0, 0.08)), sig=float(rs.uniform(0.05, 0.6)), drift=float(rs.choice([0.0, 0.01, -0.02])))
        M = int(rs.choice([2, 3, 4, 5, 6]))
        l = int(rs.randint(0, 5 if M <= 3 else 4))
        n = int(rs.choice([1, 7, 33, 129, 257]))
        Z = rs.randn(M ** l, n)
        variant = case % 3
        for cname, pay in (("Euro_GBM", O.EURO), ("Asian_GBM", O.ASIAN), ("Lookback_GBM", O.LOOKBACK), ("Digital_GBM", O.DIGITAL)):
            if variant == 2 and (pay not in (O.EURO, O.ASIAN) or M % 2):
                continue
            scheme = "milstein" if variant == 1 else "em"
            p = O.Params(scheme=scheme, **kw)
            opt = getattr(mm, cname)(scheme=scheme, **kw)
            if variant == 2:
                rf, rc = c_oracle.gbm_anti_replay(p, pay, l, M, Z)
                Pf, Pc = opt.payoff_replay(Z, l, M, arith="exact", anti=True)
                Qf, Qc = opt.payoff_replay(Z, l, M, arith="fast", anti=True)
            else:
                rf, rc = c_oracle.gbm_replay(p, pay, l, M, Z)
                Pf, Pc = opt.payoff_replay(Z, l, M, arith="exact")
                Qf, Qc = opt.payoff_replay(Z, l, M, arith="fast")
            assert np.array_equal(Pf, rf) and np.array_equal(Pc, rc), (case, cname, kw, M, l, n)
            if pay != O.DIGITAL:
                assert np.all(np.abs(Qf - rf) <= payoff_tol(p, rf)) and np.all(np.abs(Qc - rc) <= payoff_tol(p, rc)), (case, cname)
            n_checked += 1
    assert n_checked > 90


def test_randomised_merton_parameters_against_numpy_oracle(cuda_lib):
    """Seeded random Merton parameter sets (jump intensity up to 6 per year, both schemes): the CUDA path
    replays the draws recorded from the pinned numpy restatement within the north-star tolerance."""
    rs = np.random.RandomState(424242)
    n_checked = 0
    for case in range(16):
        kw = dict(X0=float(rs.uniform(60, 140)), K=float(rs.uniform(70, 130)), T=float(rs.choice([0.5, 1.0, 2.0])),
                  r=float(rs.uniform(0.0, 0.06)), sig=float(rs.uniform(0.1, 0.5)), lam=float(rs.choice([0.5, 1.0, 3.0, 6.0])),
                  jumpmean=float(rs.uniform(-0.2, 0.2)), jumpstd=float(rs.uniform(0.05, 0.4)))
        M = int(rs.choice([2, 3, 4]))
        l = int(rs.randint(0, 4))
        scheme = "milstein" if case % 2 else "em"
        for cname, pay in (("Euro_Merton", O.EURO), ("Asian_Merton", O.ASIAN), ("Lookback_Merton", O.LOOKBACK)):
            p = O.Params(scheme=scheme, **kw)
            rec = O.RecordingDraws()
            np.random.seed(1000 + case)
            rf, rc = O.level_payoffs(p, O.MERTON, pay, l, M, n=33, draws=rec)
            opt = getattr(mm, cname)(scheme=scheme, **kw)
            for arith in ("exact", "fast"):
                Pf, Pc = opt.payoff_replay(rec.packed(), l, M, arith=arith)
                assert np.all(np.abs(Pf - rf) <= payoff_tol(p, rf)), (case, cname, arith, np.abs(Pf - rf).max())
                assert np.all(np.abs(Pc - np.asarray(rc, dtype=float)) <= payoff_tol(p, np.asarray(rc, dtype=float))), (case, cname, arith)
            n_checked += 1
    assert n_checked == 48
