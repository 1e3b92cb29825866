"""CPU: the restatements in oracle/ against fixtures generated from the reference itself."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import amer_cases, case_kwargs, payoff_tol
from tests.helpers import (CLASS_CODES, anti_cases, gbm_cases, merton_cases, milstein_gbm_cases,
                           milstein_merton_cases, oracle_params, regen_Z)


def test_np_oracle_gbm_bit_exact(golden):
    n = 0
    for c in gbm_cases(golden["payoffs_gbm"]):
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = O.level_payoffs(p, model, pay, c["l"], c["M"], Z=regen_Z(c).copy())
        assert np.array_equal(Pf, c["Pf"]), c
        assert np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"]), c
        n += 1
    assert n > 100


def test_np_oracle_gbm_global_rng_bit_exact(golden):
    # consuming numpy's global stream in the reference's order gives the same result as replay
    for c in list(gbm_cases(golden["payoffs_gbm"]))[::7]:
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        np.random.seed(c["seed"])
        Pf, _ = O.level_payoffs(p, model, pay, c["l"], c["M"], n=c["n"])
        assert np.array_equal(Pf, c["Pf"])


def test_np_oracle_merton_bit_exact(golden):
    n = 0
    for c in merton_cases(golden["payoffs_merton"]):
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = O.level_payoffs(p, model, pay, c["l"], c["M"], n=c["n"], draws=O.ReplayDraws(c["packed"]))
        assert np.array_equal(Pf, c["Pf"]), c["idx"]
        assert np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"]), c["idx"]
        # and from the seed alone, consuming the global stream in the reference's order
        if n % 5 == 0:
            np.random.seed(c["seed"])
            Pf2, _ = O.level_payoffs(p, model, pay, c["l"], c["M"], n=c["n"])
            assert np.array_equal(Pf2, c["Pf"])
        n += 1
    assert n > 60


def test_np_oracle_looper_bit_exact(golden):
    g = golden["looper"]
    for i in range(int(g["n_cases"])):
        cname, l, M, Nl, Npl, seed = [str(x) for x in g[f"c{i}_meta"]]
        model, pay = CLASS_CODES[cname]
        np.random.seed(int(seed))
        s = O.looper(O.Params(), model, pay, int(Nl), int(l), int(M), Npl=int(Npl))
        assert np.array_equal(s, g[f"c{i}_sums"]), cname


def test_c_oracle_gbm_bit_exact(golden, c_oracle):
    for c in gbm_cases(golden["payoffs_gbm"]):
        _, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = c_oracle.gbm_replay(p, pay, c["l"], c["M"], regen_Z(c))
        assert np.array_equal(Pf, c["Pf"]), c["idx"]
        assert np.array_equal(Pc, c["Pc"]), c["idx"]


def test_c_oracle_merton_close(golden, c_oracle):
    # libm exp vs numpy exp may differ in the last bit -> a few ulp on S after a jump
    for c in merton_cases(golden["payoffs_merton"]):
        _, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = c_oracle.merton_replay(p, pay, c["l"], c["M"], c["packed"])
        np.testing.assert_allclose(Pf, c["Pf"], rtol=0, atol=2e-13 * p.X0)
        np.testing.assert_allclose(Pc, c["Pc"], rtol=0, atol=2e-13 * p.X0)


def test_seven_sum_identity(golden):
    # sum dP^2 = sum Pf^2 - 2 sum Pf Pc + sum Pc^2  
    g = golden["looper"]
    for i in range(int(g["n_cases"])):
        s = g[f"c{i}_sums"]
        if int(g[f"c{i}_meta"][1]) == 0:
            assert s[0] == s[2] and s[1] == s[3] and not s[4:].any()
        else:
            assert s[1] == pytest.approx(s[3] - 2 * s[6] + s[5], rel=1e-9, abs=1e-9 * s[3])


def test_c_oracle_rng_mode_matches_closed_form(c_oracle):
    # the C port with its own generator prices a European call at l=0..: sanity of the statistical twin
    p = O.Params()
    n = 400000
    tot = 0.0
    for l in range(0, 5):
        s = c_oracle.level_sums(p, O.GBM, O.EURO, l, 2, n, seed=3, nthreads=4)
        tot += s[0] / n
    assert abs(tot - 10.450583572185565) < 0.25   # bias at l=4 + sampling noise


def test_oracles_antithetic_bit_exact(golden, c_oracle):
    """anti_EA_payoff / diffusion_anti_path(_avg) :162-317: numpy and C restatements vs the reference."""
    g = golden["payoffs_anti"]
    n = 0
    for c in anti_cases(g):
        _, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Z = regen_Z(c)
        Pf, Pc = O.anti_level_payoffs(p, pay, c["l"], c["M"], Z=Z.copy())
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"])
        Pf, Pc = c_oracle.gbm_anti_replay(p, pay, c["l"], c["M"], Z)
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), c["idx"]
        n += 1
    assert n > 40
    for j in range(int(g["n_looper"])):
        cname, l, M, Nl, Npl, seed = [str(x) for x in g[f"l{j}_meta"]]
        _, pay = CLASS_CODES[cname]
        np.random.seed(int(seed))
        s = O.looper(O.Params(), O.GBM, pay, int(Nl), int(l), int(M), Npl=int(Npl), anti=True)
        assert np.array_equal(s, g[f"l{j}_sums"])
    with pytest.raises(Exception, match="odd coarseness factor"):
        O.anti_level_payoffs(O.Params(), O.EURO, 2, 3, n=4)
    assert "odd coarseness factor" in str(g["odd_M_message"])


def test_oracles_antithetic_with_milstein_stepper(golden, c_oracle):
    """anti_payoff with milstein_GBM bound over sde (diffusion_anti_path calls self.sde, :200-210)."""
    g = golden["payoffs_anti_milstein"]
    n = 0
    for c in anti_cases(g):
        _, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"], scheme="milstein")
        Z = regen_Z(c)
        Pf, Pc = O.anti_level_payoffs(p, pay, c["l"], c["M"], Z=Z.copy())
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"])
        Pf, Pc = c_oracle.gbm_anti_replay(p, pay, c["l"], c["M"], Z)
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), c["idx"]
        n += 1
    assert n > 40
    cname, l, M, Nl, Npl, seed = [str(x) for x in g["l0_meta"]]
    np.random.seed(int(seed))
    s = O.looper(O.Params(scheme="milstein"), O.GBM, CLASS_CODES[cname][1], int(Nl), int(l), int(M), Npl=int(Npl), anti=True)
    assert np.array_equal(s, g["l0_sums"])


def test_oracles_merton_deep_levels(golden, c_oracle):
    """128 / 256 fine steps (l = 7, 8 at M = 2; l = 4 at M = 4) and 512 / 1024 (l = 9 at M = 2; l = 5 at M = 4):
    numpy restatement bit-exact, C port to libm's exp."""
    n = 0
    for c in merton_cases(golden["payoffs_merton_deep"]):
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"])
        Pf, Pc = O.level_payoffs(p, model, pay, c["l"], c["M"], n=c["n"], draws=O.ReplayDraws(c["packed"]))
        assert np.array_equal(Pf, c["Pf"]), c["idx"]
        assert np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"]), c["idx"]
        Pf, Pc = c_oracle.merton_replay(p, pay, c["l"], c["M"], c["packed"])
        np.testing.assert_allclose(Pf, c["Pf"], rtol=0, atol=2e-13 * p.X0)
        np.testing.assert_allclose(Pc, c["Pc"], rtol=0, atol=2e-13 * p.X0)
        n += 1
    assert n == 28


def test_oracles_milstein(golden, c_oracle):
    """The notebook's Milstein time-steppers bound over the reference's sde (ExampleUsages.ipynb cells 1, 6)."""
    g = golden["payoffs_milstein"]
    n = 0
    for c in milstein_gbm_cases(g):
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"], scheme="milstein")
        Z = regen_Z(c)
        Pf, Pc = O.level_payoffs(p, model, pay, c["l"], c["M"], Z=Z.copy())
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(np.asarray(Pc, dtype=float) * np.ones(c["n"]), c["Pc"])
        Pf, Pc = c_oracle.gbm_replay(p, pay, c["l"], c["M"], Z)
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(Pc, c["Pc"]), c["idx"]
        n += 1
    for c in milstein_merton_cases(g):
        model, pay = CLASS_CODES[c["cname"]]
        p = oracle_params(c["cname"], c["kw_name"], scheme="milstein")
        Pf, Pc = O.level_payoffs(p, model, pay, c["l"], c["M"], n=c["n"], draws=O.ReplayDraws(c["packed"]))
        assert np.array_equal(Pf, c["Pf"]), c["idx"]
        Pf, Pc = c_oracle.merton_replay(p, pay, c["l"], c["M"], c["packed"])
        np.testing.assert_allclose(Pf, c["Pf"], rtol=0, atol=2e-13 * p.X0)
        np.testing.assert_allclose(Pc, c["Pc"], rtol=0, atol=2e-13 * p.X0)
        n += 1
    assert n > 100


def test_oracles_american_put(golden, c_oracle):
    """Amer_GBM.path :1511-1615 + Amer_payoff :138-158 with the notebook's exercise boundary."""
    g = golden["payoffs_amer"]
    n = 0
    for c in amer_cases(g):
        p = O.Params(boundary_c=c["c"], **case_kwargs("Euro_GBM", c["kw_name"]))
        Z = regen_Z(c)
        Pf, Pc = O.level_payoffs(p, O.GBM, O.AMERICAN, c["l"], 2, Z=Z.copy())
        assert np.array_equal(Pf, c["Pf"]) and np.array_equal(np.asarray(Pc) * np.ones(c["n"]), c["Pc"])
        Pf, Pc = c_oracle.gbm_replay(p, O.AMERICAN, c["l"], 2, Z)          # libm exp vs numpy exp: last-bit
        np.testing.assert_allclose(Pf, c["Pf"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(Pc, c["Pc"], rtol=0, atol=1e-12)
        n += 1
    assert n >= 24
    np.random.seed(9600)
    assert np.array_equal(O.looper(O.Params(boundary_c=0.5), O.GBM, O.AMERICAN, 700, 3, 2, Npl=300), g["looper_sums"])
    with pytest.raises(ValueError, match="M must be equal to 2"):
        O.level_payoffs(O.Params(), O.GBM, O.AMERICAN, 2, 4, n=4)
    assert "M must be equal to 2" in str(g["bad_M_message"])


def _factor_walk(p, functional, l, M, packed):
    """The kernel's factor form of the jump-adapted scheme (csrc/mlmcb_kernels.cuh: FactorWalk) in plain
    Python, on the reference's own draws.  A jump is the record (t1, J, rho, kappa[, d1, d2]) computed from the draws
    alone; the walk multiplies: S_f (1 + t1) J at a jump, S_f (1 + kappa + rho t) to close a step (t = a + b z is the
    factor the sampler delivers for the whole step), and the coarse path takes the SUM of the fine t's."""
    N = M ** l
    T, X0 = p.T, p.X0
    mu = p.r - p.lam * p.J_bar
    dt = T / N
    sqrt_dt = np.sqrt(dt)
    a, b = mu * dt, p.sig * sqrt_dt
    n = len(packed["offW"]) - 1
    out = [np.zeros(n) for _ in range(4)]
    for i in range(n):
        zW = packed["zW"][packed["offW"][i]:packed["offW"][i + 1]]
        zQ = packed["zQ"][packed["offQ"][i]:packed["offQ"][i + 1]]
        E = packed["E"][packed["offE"][i]:packed["offE"][i + 1]]
        # ---- records: everything a jump does, before the walk (arrival times, then one record per arrival) ----
        iw = iq = 0
        tau = E[0]
        ie = 1
        recs = {}                 # step -> list of records
        grid_z = {}
        tprev, jprev = 0.0, 0
        for j in range(1, N + 1):
            tn = j * T / N                                          # :456
            while tau < tn:                                         # :457
                d1 = tau - (tprev if jprev == j else (j - 1) * T / N if j > 1 else 0.0)
                d2 = tn - tau
                t1 = mu * d1 + p.sig * (zW[iw] * np.sqrt(d1))
                J = np.exp(p.jumpmean + p.jumpstd * zQ[iq])
                rho = np.sqrt(d2) / sqrt_dt
                recs.setdefault(j, []).append((t1, J, rho, mu * d2 - rho * a, d1, d2))
                iw += 1
                iq += 1
                tprev, jprev = tau, j
                tau += E[ie]
                ie += 1
            grid_z[j] = zW[iw]
            iw += 1
        # ---- walk ----
        Sf = Sc = X0
        Gf = Gc = X0 if functional == O.LOOKBACK else 0.0
        tsum = dtc = 0.0
        for j in range(1, N + 1):
            rho, kappa, h = 1.0, 0.0, dt
            for (t1, J, r_, k_, d1, d2) in recs.get(j, ()):
                tsum += t1
                if functional == O.ASIAN:
                    dtc += d1
                    uf, uc = Sf + Sf * t1, Sc + Sc * tsum
                    Gf += 0.5 * d1 * (Sf + uf)
                    Gc += 0.5 * dtc * (Sc + uc)
                    Sf, Sc, dtc = uf * J, uc * J, 0.0
                else:
                    Sf, Sc = (Sf + Sf * t1) * J, (Sc + Sc * tsum) * J
                    if functional == O.LOOKBACK:
                        Gf, Gc = min(Gf, Sf), min(Gc, Sc)
                tsum = 0.0
                rho, kappa, h = r_, k_, d2
            tt = kappa + rho * (a + b * grid_z[j])
            uf = Sf + Sf * tt
            if functional == O.ASIAN:
                Gf += 0.5 * h * (Sf + uf)
                dtc += h
            Sf = uf
            if functional == O.LOOKBACK:
                Gf = min(Gf, Sf)
            tsum += tt
            if j % M == 0:
                uc = Sc + Sc * tsum
                if functional == O.ASIAN:
                    Gc += 0.5 * dtc * (Sc + uc)
                    dtc = 0.0
                Sc = uc
                if functional == O.LOOKBACK:
                    Gc = min(Gc, Sc)
                tsum = 0.0
        out[0][i], out[1][i], out[2][i], out[3][i] = Sf, Gf, Sc, Gc
    if functional == O.ASIAN:
        return out[1] / T, out[3] / T
    if functional == O.LOOKBACK:
        return out[0], out[1], out[2], out[3]
    return out[0], out[2]


def test_factor_form_of_the_jump_adapted_scheme_reproduces_the_reference(golden):
    """The algebra the Merton factor walk rests on - jumps as state-independent factors, the closing step affine in the
    whole-step factor, the coarse increment as the sum of the fine ones - against the REFERENCE's per-path payoffs on
    its own draws (fixtures made by oracle/make_golden.py from /root/reference): 1e-12 like the replay kernels, for the
    four products at 1 .. 1024 fine steps, M = 2, 3 and 4."""
    n = 0
    for name in ("payoffs_merton", "payoffs_merton_deep"):
        for c in merton_cases(golden[name]):
            if c["idx"] % 2 and name == "payoffs_merton":
                continue                                            # every other shallow case keeps the test quick
            model, pay = CLASS_CODES[c["cname"]]
            p = oracle_params(c["cname"], c["kw_name"])
            path_out = _factor_walk(p, O.functional_of(pay), c["l"], c["M"], c["packed"])
            Pf, Pc = O.apply_payoff(p, pay, c["l"], c["M"], path_out)
            if c["l"] == 0:
                Pc = c["Pc"]                                        # l = 0: the raw coarse output is not a payoff (:82)
            if c["cname"] == "Digital_Merton":
                assert np.array_equal(Pf, c["Pf"]) and np.array_equal(np.asarray(Pc, dtype=float), c["Pc"]), c["idx"]
            else:
                assert np.all(np.abs(Pf - c["Pf"]) <= payoff_tol(p, c["Pf"])), (name, c["idx"], np.abs(Pf - c["Pf"]).max())
                assert np.all(np.abs(Pc - c["Pc"]) <= payoff_tol(p, c["Pc"])), (name, c["idx"], np.abs(Pc - c["Pc"]).max())
            n += 1
    assert n >= 60
