"""Shared test helpers: golden-case decoding and oracle parameter construction."""
import numpy as np

from oracle import np_oracle as O

ALT = {
    "default": {},
    "alt": dict(X0=125, K=105, r=0.02, sig=0.5, T=1),
    "alt2": dict(X0=90, K=100, r=0.03, sig=0.35, T=2.0, drift=0.01),
    "alt_mjd": dict(X0=110, K=100, r=0.03, sig=0.3, T=0.5, lam=3, jumpmean=-0.05, jumpstd=0.3),
}
CLASS_CODES = {
    "Euro_GBM": (O.GBM, O.EURO), "Asian_GBM": (O.GBM, O.ASIAN), "Lookback_GBM": (O.GBM, O.LOOKBACK),
    "Digital_GBM": (O.GBM, O.DIGITAL), "Euro_Merton": (O.MERTON, O.EURO), "Asian_Merton": (O.MERTON, O.ASIAN),
    "Lookback_Merton": (O.MERTON, O.LOOKBACK), "Digital_Merton": (O.MERTON, O.DIGITAL),
}


def case_kwargs(cname, kw_name):
    if kw_name == "alt" and cname.endswith("Merton"):
        kw = dict(ALT["alt_mjd"])
    else:
        kw = dict(ALT[kw_name])
    if cname.startswith("Lookback"):
        kw.pop("K", None)
    return kw


def gbm_cases(g):
    for i in range(int(g["n_cases"])):
        cname, kw_name, l, M, n, seed = [str(x) for x in g[f"c{i}_meta"]]
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed),
                   Pf=g[f"c{i}_Pf"], Pc=g[f"c{i}_Pc"], Z=g[f"c{i}_Z"] if f"c{i}_Z" in g.files else None)


def regen_Z(case):
    """The normals the reference consumed: legacy RandomState streams are frozen (NEP 19)."""
    if case["Z"] is not None:
        return case["Z"]
    rs = np.random.RandomState(case["seed"])
    return np.stack([rs.randn(case["n"]) for _ in range(case["M"] ** case["l"])])


def merton_cases(g):
    for i in range(int(g["n_cases"])):
        cname, kw_name, l, M, n, seed = [str(x) for x in g[f"c{i}_meta"]]
        packed = dict(zW=g[f"c{i}_zW"], offW=g[f"c{i}_offW"], zQ=g[f"c{i}_zQ"], offQ=g[f"c{i}_offQ"],
                      E=g[f"c{i}_zE"], offE=g[f"c{i}_offE"])
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed),
                   Pf=g[f"c{i}_Pf"], Pc=g[f"c{i}_Pc"], packed=packed)


def oracle_params(cname, kw_name, scheme="em"):
    return O.Params(scheme=scheme, **case_kwargs(cname, kw_name))


def payoff_tol(p, P):
    """North-star replay tolerance: |dP| <= 1e-12 * max(|P|, K e^{-rT});
    X0 stands in for K on lookbacks."""
    scale = (p.K if p.K is not None else p.X0) * np.exp(-p.r * p.T)
    return 1e-12 * np.maximum(np.abs(P), scale)


def anti_cases(g):
    for i in range(int(g["n_cases"])):
        cname, kw_name, l, M, n, seed = [str(x) for x in g[f"c{i}_meta"]]
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed),
                   Pf=g[f"c{i}_Pf"], Pc=g[f"c{i}_Pc"], Z=None)


def milstein_gbm_cases(g):
    for i in range(int(g["n_gbm"])):
        cname, kw_name, l, M, n, seed = [str(x) for x in g[f"g{i}_meta"]]
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed),
                   Pf=g[f"g{i}_Pf"], Pc=g[f"g{i}_Pc"], Z=None)


def milstein_merton_cases(g):
    for i in range(int(g["n_merton"])):
        cname, kw_name, l, M, n, seed = [str(x) for x in g[f"m{i}_meta"]]
        packed = dict(zW=g[f"m{i}_zW"], offW=g[f"m{i}_offW"], zQ=g[f"m{i}_zQ"], offQ=g[f"m{i}_offQ"],
                      E=g[f"m{i}_zE"], offE=g[f"m{i}_offE"])
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed),
                   Pf=g[f"m{i}_Pf"], Pc=g[f"m{i}_Pc"], packed=packed)


def amer_cases(g):
    for i in range(int(g["n_cases"])):
        cname, kw_name, l, M, n, seed, c = [str(x) for x in g[f"c{i}_meta"]]
        yield dict(idx=i, cname=cname, kw_name=kw_name, l=int(l), M=int(M), n=int(n), seed=int(seed), c=float(c),
                   Pf=g[f"c{i}_Pf"], Pc=g[f"c{i}_Pc"], Z=None)
