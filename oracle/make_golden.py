"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  Every number written here comes out of /root/reference/MLMC_RSCAM.py
itself (imported through oracle/ref_loader.py); the restatements in oracle/ are not involved.

Files
  payoffs_gbm.npz     per-path (Pf,Pc) of `opt.payoff(N,l,M)` for the 4 GBM products, with the
                      seed; small cases also carry the literal normals Z[step,path] the
                      reference consumed (np.random.randn wrapped while the reference ran)
  payoffs_anti_milstein.npz  anti_payoff with the notebook's Milstein stepper bound over `sde`
  payoffs_merton_deep.npz    the 4 Merton products at 128 / 256 fine steps (l=7,8 M=2; l=4 M=4) and at 512 / 1024 (l=9 M=2; l=5 M=4)
  payoffs_merton.npz  same for the 4 Merton products with the three ragged draw streams
                      (dW normals, jump-size normals, inter-arrival times) in CSR layout
  looper.npz          `opt.looper(Nl,l,M,False,Npl)` 7-vectors for seeds
  mlmc_trace.npz      a full `opt.mlmc(eps)` run: every looper request (Nl,l) and the 7 sums it
                      returned, plus the final (sums,N,alpha,beta) -> pins the host driver logic
  closed_forms.npz    BS() values
"""
import io
import os
import sys
import contextlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GBM_CLASSES = ["Euro_GBM", "Asian_GBM", "Lookback_GBM", "Digital_GBM"]
MJD_CLASSES = ["Euro_Merton", "Asian_Merton", "Lookback_Merton", "Digital_Merton"]
# non-default parameter sets so every field of the parameter struct is exercised
ALT_GBM = dict(X0=125, K=105, r=0.02, sig=0.5, T=1)
ALT_GBM2 = dict(X0=90, K=100, r=0.03, sig=0.35, T=2.0, drift=0.01)
ALT_MJD = dict(X0=110, K=100, r=0.03, sig=0.3, T=0.5, lam=3, jumpmean=-0.05, jumpstd=0.3)


class Tap:
    """Wraps numpy's legacy global draw functions while the reference runs, logging values."""

    def __init__(self):
        self.randn_calls, self.stdn, self.expo = [], [], []

    def __enter__(self):
        self._randn, self._stdn = np.random.randn, np.random.standard_normal
        tap = self

        def randn(*a):
            v = tap._randn(*a)
            tap.randn_calls.append(np.array(v, dtype=np.float64, copy=True))
            return v

        def stdn(*a, **k):
            v = tap._stdn(*a, **k)
            tap.stdn.append(float(v))
            return v

        np.random.randn, np.random.standard_normal = randn, stdn
        return self

    def __exit__(self, *exc):
        np.random.randn, np.random.standard_normal = self._randn, self._stdn

    def wrap_jumptime(self, opt):
        inner = opt.jumptime  # np.random.exponential bound at construction (MLMC_RSCAM.py:1294)

        def jumptime(scale=1.0):
            v = inner(scale=scale)
            self.expo.append(float(v))
            return v

        opt.jumptime = jumptime


def gbm_cases(ref):
    out = {}
    idx = 0
    for cname in GBM_CLASSES:
        for kw_name, kw in (("default", {}), ("alt", ALT_GBM), ("alt2", ALT_GBM2)):
            for M in (2, 4, 3):
                for l in (0, 1, 2, 3, 5, 6):
                    if kw_name != "default" and l in (5,):
                        continue
                    if M == 3 and (l > 3 or kw_name != "default"):
                        continue
                    steps = M ** l
                    n = 128 if steps <= 64 else 48
                    seed = 1000 + idx
                    kw2 = dict(kw)
                    if cname == "Lookback_GBM":
                        kw2.pop("K", None)
                    opt = getattr(ref, cname)(**kw2)
                    np.random.seed(seed)
                    with Tap() as tap:
                        Pf, Pc = opt.payoff(n, l, M)
                    Z = np.stack(tap.randn_calls)          # (steps, n) in call order
                    assert Z.shape == (steps, n)
                    key = f"c{idx}"
                    out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[key + "_Pf"] = np.asarray(Pf, dtype=np.float64)
                    out[key + "_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                    if steps * n <= 8192 and kw_name == "default":
                        out[key + "_Z"] = Z
                    # the seed alone regenerates Z: legacy RandomState streams are frozen (NEP 19)
                    np.random.seed(seed)
                    Z2 = np.stack([np.random.randn(n) for _ in range(steps)])
                    assert np.array_equal(Z, Z2)
                    idx += 1
    out["n_cases"] = np.array(idx)
    return out


def merton_cases(ref):
    out = {}
    idx = 0
    for cname in MJD_CLASSES:
        for kw_name, kw in (("default", {}), ("alt", ALT_MJD)):
            for M in (2, 4, 3):
                for l in (0, 1, 2, 3, 5):
                    if M == 3 and (l > 2 or kw_name != "default"):
                        continue
                    if l == 5 and M == 4:
                        continue
                    n = 64 if M ** l <= 16 else 24
                    seed = 5000 + idx
                    kw2 = dict(kw)
                    if cname == "Lookback_Merton":
                        kw2.pop("K", None)
                    opt = getattr(ref, cname)(**kw2)
                    np.random.seed(seed)
                    Pf_all, Pc_all = [], []
                    zw, zq, ex = [], [], []
                    for _ in range(n):      # one path per call == one N-path call (same global stream)
                        with Tap() as tap:
                            tap.wrap_jumptime(opt)
                            Pf, Pc = opt.payoff(1, l, M)
                            opt.jumptime = np.random.exponential
                        Pf_all.append(float(Pf[0]))
                        Pc_all.append(float(np.asarray(Pc).reshape(-1)[0]))
                        zw.append([float(v) for v in tap.randn_calls])
                        zq.append(tap.stdn)
                        ex.append(tap.expo)
                    # cross-check: one N-path call from the same seed gives the same payoffs
                    opt2 = getattr(ref, cname)(**kw2)
                    np.random.seed(seed)
                    Pf2, Pc2 = opt2.payoff(n, l, M)
                    assert np.array_equal(np.array(Pf_all), Pf2)

                    def csr(rows):
                        off = np.zeros(len(rows) + 1, dtype=np.int64)
                        off[1:] = np.cumsum([len(r) for r in rows])
                        return np.array([v for r in rows for v in r], dtype=np.float64), off

                    key = f"c{idx}"
                    out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[key + "_Pf"] = np.array(Pf_all)
                    out[key + "_Pc"] = np.array(Pc_all)
                    for nm, rows in (("W", zw), ("Q", zq), ("E", ex)):
                        flat, off = csr(rows)
                        out[key + "_z" + nm] = flat
                        out[key + "_off" + nm] = off
                    idx += 1
    out["n_cases"] = np.array(idx)
    return out


def anti_cases(ref):
    """anti_payoff (:292-317) per-path outputs for Euro_GBM / Asian_GBM, even M."""
    out = {}
    idx = 0
    for cname in ("Euro_GBM", "Asian_GBM"):
        for kw_name, kw in (("default", {}), ("alt2", ALT_GBM2)):
            for M in (2, 4, 6):
                for l in (0, 1, 2, 3, 5):
                    if M == 6 and l > 3:
                        continue
                    steps = M ** l
                    n = 96 if steps <= 64 else 32
                    seed = 7000 + idx
                    opt = getattr(ref, cname)(**kw)
                    np.random.seed(seed)
                    with Tap() as tap:
                        Pf, Pc = opt.anti_payoff(n, l, M)
                    Z = np.stack(tap.randn_calls)
                    assert Z.shape == (steps, n)
                    key = f"c{idx}"
                    out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[key + "_Pf"] = np.asarray(Pf, dtype=np.float64)
                    out[key + "_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                    idx += 1
    # looper(anti=True) sums and the odd-M exception text
    for j, (cname, l, M, Nl, Npl) in enumerate((("Euro_GBM", 2, 2, 1500, 1000), ("Asian_GBM", 3, 4, 700, 300),
                                                 ("Asian_GBM", 0, 2, 900, 400))):
        opt = getattr(ref, cname)()
        np.random.seed(7700 + j)
        out[f"l{j}_meta"] = np.array([cname, str(l), str(M), str(Nl), str(Npl), str(7700 + j)])
        out[f"l{j}_sums"] = opt.looper(Nl, l, M, True, Npl)
    try:
        ref.Euro_GBM().looper(10, 2, 3, True)
    except Exception as e:  # noqa: BLE001
        out["odd_M_message"] = np.array(str(e))
    out["n_cases"] = np.array(idx)
    out["n_looper"] = np.array(3)
    return out


# The demo notebook (ExampleUsages.ipynb cell 1) swaps the time-stepper by binding these two
# functions over `opt.sde` with types.MethodType (cell 6).  They are the specification of the
# Milstein option of the new backend, so they are restated here to drive the unmodified reference.
def milstein_GBM(self, X, dt, t_, dW):
    return self.mu * X * dt + self.sig * X * dW + 0.5 * X * (self.sig ** 2) * (dW ** 2 - dt)


def milstein_Merton(self, X, dt, t_, dW, dJ=0):
    dx_ = (self.r - self.lam * self.J_bar) * X * dt + self.sig * X * dW + 0.5 * X * (self.sig ** 2) * (dW ** 2 - dt)
    return dx_ + (dx_ + X) * dJ


def milstein_cases(ref):
    import types
    out = {}
    idx = 0
    for cname in GBM_CLASSES:
        for kw_name, kw in (("default", {}), ("alt2", ALT_GBM2)):
            for M in (2, 4, 3):
                for l in (0, 1, 2, 3, 5):
                    if M == 3 and l > 2:
                        continue
                    steps = M ** l
                    n = 96 if steps <= 64 else 32
                    seed = 8000 + idx
                    kw2 = dict(kw)
                    opt = getattr(ref, cname)(**kw2)
                    opt.sde = types.MethodType(milstein_GBM, opt)
                    np.random.seed(seed)
                    Pf, Pc = opt.payoff(n, l, M)
                    out[f"g{idx}_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[f"g{idx}_Pf"] = np.asarray(Pf, dtype=np.float64)
                    out[f"g{idx}_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                    idx += 1
    out["n_gbm"] = np.array(idx)
    idx = 0
    for cname in MJD_CLASSES:
        for kw_name, kw in (("default", {}), ("alt", ALT_MJD)):
            for M in (2, 4):
                for l in (0, 1, 2, 3):
                    n = 32
                    seed = 8500 + idx
                    kw2 = dict(kw)
                    opt = getattr(ref, cname)(**kw2)
                    opt.sde = types.MethodType(milstein_Merton, opt)
                    np.random.seed(seed)
                    Pf_all, Pc_all, zw, zq, ex = [], [], [], [], []
                    for _ in range(n):
                        with Tap() as tap:
                            tap.wrap_jumptime(opt)
                            Pf, Pc = opt.payoff(1, l, M)
                            opt.jumptime = np.random.exponential
                        Pf_all.append(float(Pf[0]))
                        Pc_all.append(float(np.asarray(Pc).reshape(-1)[0]))
                        zw.append([float(v) for v in tap.randn_calls])
                        zq.append(tap.stdn)
                        ex.append(tap.expo)

                    def csr(rows):
                        off = np.zeros(len(rows) + 1, dtype=np.int64)
                        off[1:] = np.cumsum([len(r) for r in rows])
                        return np.array([v for r in rows for v in r], dtype=np.float64), off

                    key = f"m{idx}"
                    out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[key + "_Pf"] = np.array(Pf_all)
                    out[key + "_Pc"] = np.array(Pc_all)
                    for nm, rows in (("W", zw), ("Q", zq), ("E", ex)):
                        flat, off = csr(rows)
                        out[key + "_z" + nm] = flat
                        out[key + "_off" + nm] = off
                    idx += 1
    out["n_merton"] = np.array(idx)
    return out


def anti_milstein_cases(ref):
    """anti_payoff (:292-317) with the notebook's milstein_GBM bound over `sde`: diffusion_anti_path(_avg)
    calls whatever `self.sde` is (:200-210), so the antithetic estimator composes with the Milstein stepper."""
    import types
    out = {}
    idx = 0
    for cname in ("Euro_GBM", "Asian_GBM"):
        for kw_name, kw in (("default", {}), ("alt2", ALT_GBM2)):
            for M in (2, 4, 6):
                for l in (0, 1, 2, 3, 5):
                    if M == 6 and l > 2:
                        continue
                    steps = M ** l
                    n = 64 if steps <= 64 else 24
                    seed = 7300 + idx
                    opt = getattr(ref, cname)(**kw)
                    opt.sde = types.MethodType(milstein_GBM, opt)
                    np.random.seed(seed)
                    Pf, Pc = opt.anti_payoff(n, l, M)
                    key = f"c{idx}"
                    out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
                    out[key + "_Pf"] = np.asarray(Pf, dtype=np.float64)
                    out[key + "_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                    idx += 1
    opt = ref.Asian_GBM()
    opt.sde = types.MethodType(milstein_GBM, opt)
    np.random.seed(7399)
    out["l0_meta"] = np.array(["Asian_GBM", "2", "4", "600", "250", "7399"])
    out["l0_sums"] = opt.looper(600, 2, 4, True, 250)
    out["n_cases"] = np.array(idx)
    out["n_looper"] = np.array(1)
    return out


def merton_deep_cases(ref):
    """JD_path* (:428-651) at 128 and 256 fine steps (l = 7, 8 with M = 2; l = 4 with M = 4) and, 8 paths each, at
    512 and 1024 (l = 9, M = 2; l = 5, M = 4) - the long-path regime.  Same layout as payoffs_merton.npz."""
    out = {}
    idx = 0
    # the 512- and 1024-step cases were appended later: they come last so that the earlier cases keep their seeds
    plan = [(cname, kw_name, kw, l, M, 24) for cname in MJD_CLASSES for kw_name, kw in (("default", {}), ("alt", ALT_MJD))
            for (l, M) in ((7, 2), (8, 2), (4, 4)) if not (kw_name == "alt" and l == 8)]
    plan += [(cname, "default", {}, l, M, 8) for cname in MJD_CLASSES for (l, M) in ((9, 2), (5, 4))]
    for cname, kw_name, kw, l, M, n in plan:
        seed = 5600 + idx
        kw2 = dict(kw)
        if cname == "Lookback_Merton":
            kw2.pop("K", None)
        opt = getattr(ref, cname)(**kw2)
        np.random.seed(seed)
        Pf_all, Pc_all, zw, zq, ex = [], [], [], [], []
        for _ in range(n):
            with Tap() as tap:
                tap.wrap_jumptime(opt)
                Pf, Pc = opt.payoff(1, l, M)
                opt.jumptime = np.random.exponential
            Pf_all.append(float(Pf[0]))
            Pc_all.append(float(np.asarray(Pc).reshape(-1)[0]))
            zw.append([float(v) for v in tap.randn_calls])
            zq.append(tap.stdn)
            ex.append(tap.expo)

        def csr(rows):
            off = np.zeros(len(rows) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(r) for r in rows])
            return np.array([v for r in rows for v in r], dtype=np.float64), off

        key = f"c{idx}"
        out[key + "_meta"] = np.array([cname, kw_name, str(l), str(M), str(n), str(seed)])
        out[key + "_Pf"] = np.array(Pf_all)
        out[key + "_Pc"] = np.array(Pc_all)
        for nm, rows in (("W", zw), ("Q", zq), ("E", ex)):
            flat, off = csr(rows)
            out[key + "_z" + nm] = flat
            out[key + "_off" + nm] = off
        idx += 1
    out["n_cases"] = np.array(idx)
    return out


def amer_cases(ref):
    """Amer_GBM (:1483-1632) with the demo notebook's exercise boundary `exb` (cell 1):
    np.maximum(self.K*(1-0.8*np.sqrt(self.T-t)),0); c = 0.8 and one other value."""
    out = {}
    idx = 0
    for c in (0.8, 0.5):
        def exb(self, t, c=c):
            return np.maximum(self.K * (1 - c * np.sqrt(self.T - t)), 0)
        for kw_name, kw in (("default", {}), ("alt2", ALT_GBM2)):
            for l in (0, 1, 2, 3, 4, 6):
                n = 96 if l <= 4 else 32
                seed = 9500 + idx
                opt = ref.Amer_GBM(exerciseBoundary=exb, **kw)
                np.random.seed(seed)
                Pf, Pc = opt.payoff(n, l, 2)
                out[f"c{idx}_meta"] = np.array(["Amer_GBM", kw_name, str(l), "2", str(n), str(seed), repr(c)])
                out[f"c{idx}_Pf"] = np.asarray(Pf, dtype=np.float64)
                out[f"c{idx}_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                idx += 1
    opt = ref.Amer_GBM(exerciseBoundary=exb)
    np.random.seed(9600)
    out["looper_sums"] = opt.looper(700, 3, 2, False, 300)
    try:
        opt.looper(10, 2, 4, False)
    except ValueError as e:
        out["bad_M_message"] = np.array(str(e))
    out["n_cases"] = np.array(idx)
    return out


def looper_cases(ref):
    out = {}
    idx = 0
    for cname in GBM_CLASSES + MJD_CLASSES:
        for (l, M, Nl, Npl) in ((0, 2, 2500, 1000), (2, 2, 1500, 1000), (3, 4, 700, 300)):
            seed = 9000 + idx
            opt = getattr(ref, cname)()
            np.random.seed(seed)
            s = opt.looper(Nl, l, M, False, Npl)
            out[f"c{idx}_meta"] = np.array([cname, str(l), str(M), str(Nl), str(Npl), str(seed)])
            out[f"c{idx}_sums"] = s
            idx += 1
    out["n_cases"] = np.array(idx)
    return out


def mlmc_traces(ref):
    """Run the reference driver with a tapped looper.  alpha_0/beta_0 cover the three branches:
    both given, both None (regression, :1007-1018; needs the None-safe max of SURVEY 0.2, so the
    reference is given 0 there and alpha_0==None is restored on the instance afterwards is NOT
    possible - instead the regression branch is pinned with alpha_0=beta_0=None on a subclass
    whose only change is the guarded `max`)."""
    out = {}
    cases = [
        ("Euro_GBM", dict(X0=125, K=105, r=0.02, sig=0.5, alpha_0=1, beta_0=1), 0.05, 2, 1000, 11),
        ("Asian_GBM", dict(alpha_0=1, beta_0=1), 0.02, 4, 1000, 12),
        ("Lookback_GBM", dict(alpha_0=1, beta_0=1), 0.05, 4, 500, 13),
        ("Digital_GBM", dict(alpha_0=1, beta_0=0.5), 0.3, 4, 1000, 14),
        ("Euro_Merton", dict(alpha_0=1, beta_0=1), 0.3, 2, 1000, 15),
        ("Asian_GBM", dict(alpha_0=1, beta_0=1, anti=True), 0.02, 4, 1000, 16),
        # regression branches: alpha_0=None with beta_0 given is legal Python 3 only if the
        # max(0,None) at :982 is bypassed; the reference raises TypeError there,
        # which test_mlmc_none_raises_in_reference pins.  Not traced here.
    ]
    for i, (cname, kw, eps, M, N0, seed) in enumerate(cases):
        kw = dict(kw)
        anti = kw.pop("anti", False)
        opt = getattr(ref, cname)(**kw)
        calls = []
        inner = opt.looper

        def looper(Nl, l, M_, anti, Npl=10 ** 4, _inner=inner, _calls=calls):
            s = _inner(Nl, l, M_, anti, Npl)
            _calls.append((Nl, l, np.array(s, copy=True)))
            return s

        opt.looper = looper
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            sums, N = opt.mlmc(eps, M=M, anti=anti, N0=N0, warm_start=True)
        out[f"t{i}_meta"] = np.array([cname, repr(kw), repr(eps), str(M), str(N0), str(seed), str(int(anti))])
        out[f"t{i}_req"] = np.array([(c[0], c[1]) for c in calls], dtype=np.int64)
        out[f"t{i}_ret"] = np.stack([c[2] for c in calls])
        out[f"t{i}_sums"] = sums
        out[f"t{i}_N"] = N
        out[f"t{i}_ab"] = np.array([opt.alpha_0, opt.beta_0], dtype=np.float64)
    out["n_cases"] = np.array(len(cases))
    return out


def closed_forms(ref):
    out = {}
    out["Euro_GBM"] = np.array(ref.Euro_GBM().BS())
    out["Euro_GBM_readme"] = np.array(ref.Euro_GBM(X0=125, K=105, r=0.02, sig=0.5).BS())
    out["Asian_GBM"] = np.array(ref.Asian_GBM().BS())
    out["Lookback_GBM"] = np.array(ref.Lookback_GBM().BS())
    out["Digital_GBM"] = np.array(ref.Digital_GBM().BS())
    out["Euro_Merton"] = np.array(ref.Euro_Merton().BS())
    out["Euro_Merton_alt"] = np.array(ref.Euro_Merton(lam=3, jumpmean=-0.05, jumpstd=0.3, sig=0.3, X0=110).BS(n=60))
    out["Euro_GBM_alt2"] = np.array(ref.Euro_GBM(**ALT_GBM2).BS())
    return out


def amer_callable_cases(ref):
    """Amer_GBM with exercise boundaries outside the square-root family (oracle/boundaries.py): what the
    host-tabulated boundary path of the library has to reproduce."""
    from oracle.boundaries import BOUNDARIES
    out = {}
    idx = 0
    for bname, fn in BOUNDARIES.items():
        for kw_name, kw in (("default", {}), ("alt2", ALT_GBM2)):
            for l in (0, 1, 2, 4, 6):
                n = 96 if l <= 4 else 32
                seed = 9700 + idx
                opt = ref.Amer_GBM(exerciseBoundary=fn, **kw)
                np.random.seed(seed)
                Pf, Pc = opt.payoff(n, l, 2)
                out[f"c{idx}_meta"] = np.array(["Amer_GBM", kw_name, str(l), "2", str(n), str(seed), bname])
                out[f"c{idx}_Pf"] = np.asarray(Pf, dtype=np.float64)
                out[f"c{idx}_Pc"] = np.asarray(Pc, dtype=np.float64) * np.ones(n)
                idx += 1
    out["n_cases"] = np.array(idx)
    return out


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("payoffs_gbm", gbm_cases), ("payoffs_merton", merton_cases), ("payoffs_anti", anti_cases),
                     ("payoffs_anti_milstein", anti_milstein_cases), ("payoffs_merton_deep", merton_deep_cases),
                     ("payoffs_milstein", milstein_cases), ("payoffs_amer", amer_cases),
                     ("payoffs_amer_callable", amer_callable_cases), ("looper", looper_cases),
                     ("mlmc_trace", mlmc_traces), ("closed_forms", closed_forms)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:       # python oracle/make_golden.py [file ...]
            continue
        d = fn(ref)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, len(d), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
