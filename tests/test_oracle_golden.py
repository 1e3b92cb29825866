"""CPU: the restatements in oracle/ against fixtures generated from the reference itself."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import amer_cases, case_kwargs
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
