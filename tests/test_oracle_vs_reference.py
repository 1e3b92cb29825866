"""CPU, build container only: oracle restatements vs the UNMODIFIED reference, same seed."""
import numpy as np
import pytest

from oracle import np_oracle as O
from oracle.ref_loader import load_reference, reference_available
from tests.helpers import CLASS_CODES

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")


@pytest.mark.parametrize("cname", sorted(CLASS_CODES))
def test_payoff_and_looper_bit_identical(cname):
    ref = load_reference()
    model, pay = CLASS_CODES[cname]
    for l, M in ((0, 2), (1, 2), (2, 4), (3, 3), (4, 2)):
        opt = getattr(ref, cname)()
        np.random.seed(17 + l)
        a = opt.payoff(40, l, M)
        np.random.seed(17 + l)
        b = O.level_payoffs(O.Params(), model, pay, l, M, n=40)
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(np.asarray(a[1], dtype=float), np.asarray(b[1], dtype=float))
        np.random.seed(5)
        s1 = opt.looper(130, l, M, False, Npl=50)
        np.random.seed(5)
        s2 = O.looper(O.Params(), model, pay, 130, l, M, Npl=50)
        assert np.array_equal(s1, s2)


def test_reference_mlmc_none_raises_typeerror():
    # the README call fails on Python 3 at MLMC_RSCAM.py:982
    ref = load_reference()
    with pytest.raises(TypeError):
        ref.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5).mlmc(eps=0.5)


def test_host_driver_equals_reference_driver_live():
    """Both drivers fed by the same reference looper on the same numpy stream -> identical output."""
    import contextlib
    import io
    from multilevel_mc_sdes_b200 import mlmc_driver
    ref = load_reference()
    for cname, kw, eps, M in (("Euro_GBM", dict(alpha_0=1, beta_0=1), 0.1, 2),
                              ("Lookback_GBM", dict(alpha_0=1, beta_0=1), 0.1, 4),
                              ("Digital_GBM", dict(alpha_0=1, beta_0=0.5), 0.5, 4)):
        opt = getattr(ref, cname)(**kw)
        np.random.seed(99)
        with contextlib.redirect_stdout(io.StringIO()):
            s_ref, N_ref = opt.mlmc(eps, M=M, warm_start=False)
        opt2 = getattr(ref, cname)(**kw)
        np.random.seed(99)
        s, N, a, b = mlmc_driver(lambda reqs, M_: [opt2.looper(n, l, M_, False) for n, l in reqs],
                                 eps, M, 1000, kw["alpha_0"], kw["beta_0"])
        assert np.array_equal(s, s_ref) and np.array_equal(N, N_ref)
