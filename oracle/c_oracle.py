"""ctypes face of oracle/c_oracle.c.  TEST INFRASTRUCTURE ONLY (see c_oracle.c header)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmlmc_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("X0", "K", "T", "r", "sig", "drift", "lam", "jumpmean", "jumpstd", "beta", "disc", "J_bar",
                 "milstein", "boundary_c")]


def build(force=False):
    src = os.path.join(_HERE, "c_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmlmc_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_gbm_replay.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp]
        L.orc_merton_replay.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, C.c_int64,
                                        dp, ip, dp, ip, dp, ip, dp, dp]
        L.orc_level_sums.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     C.c_uint64, C.c_int, dp]
        L.orc_gbm_anti_replay.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp]
        L.orc_level_sums_anti.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, C.c_int64,
                                          C.c_uint64, C.c_int, dp]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def to_params(p):
    """p: oracle.np_oracle.Params (or anything with the same attributes)."""
    return OrcParams(float(p.X0), float(p.K if p.K is not None else 0.0), float(p.T), float(p.r), float(p.sig),
                     float(getattr(p, "drift", 0.0)), float(getattr(p, "lam", 1.0)),
                     float(getattr(p, "jumpmean", 0.0)), float(getattr(p, "jumpstd", 0.0)),
                     float(getattr(p, "beta", 0.5826)), float(np.exp(-p.r * p.T)),
                     float(np.exp(getattr(p, "jumpmean", 0.0) + 0.5 * getattr(p, "jumpstd", 0.0) ** 2) - 1),
                     1.0 if getattr(p, "scheme", "em") == "milstein" else 0.0, float(getattr(p, "boundary_c", 0.8)))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def gbm_replay(p, payoff, l, M, Z):
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    n = Z.shape[1]
    assert Z.shape[0] == M ** l
    Pf, Pc = np.empty(n), np.empty(n)
    cp = to_params(p)
    lib().orc_gbm_replay(C.byref(cp), payoff, l, M, n, _dp(Z), _dp(Pf), _dp(Pc))
    return Pf, Pc


def gbm_anti_replay(p, payoff, l, M, Z):
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    n = Z.shape[1]
    assert Z.shape[0] == M ** l
    Pf, Pc = np.empty(n), np.empty(n)
    cp = to_params(p)
    rc = lib().orc_gbm_anti_replay(C.byref(cp), payoff, l, M, n, _dp(Z), _dp(Pf), _dp(Pc))
    assert rc == 0
    return Pf, Pc


def merton_replay(p, payoff, l, M, packed):
    n = len(packed["offW"]) - 1
    Pf, Pc = np.empty(n), np.empty(n)
    cp = to_params(p)
    a = {k: np.ascontiguousarray(v) for k, v in packed.items()}
    lib().orc_merton_replay(C.byref(cp), payoff, l, M, n, _dp(a["zW"]), _ip(a["offW"]), _dp(a["zQ"]),
                            _ip(a["offQ"]), _dp(a["E"]), _ip(a["offE"]), _dp(Pf), _dp(Pc))
    return Pf, Pc


def level_sums(p, model, payoff, l, M, n, seed=0, nthreads=1):
    out = np.zeros(7)
    cp = to_params(p)
    lib().orc_level_sums(C.byref(cp), model, payoff, l, M, int(n), int(seed), int(nthreads), _dp(out))
    return out


def level_sums_anti(p, payoff, l, M, n, seed=0, nthreads=1):
    out = np.zeros(7)
    cp = to_params(p)
    rc = lib().orc_level_sums_anti(C.byref(cp), payoff, l, M, int(n), int(seed), int(nthreads), _dp(out))
    assert rc == 0
    return out


def max_threads():
    return int(lib().orc_max_threads())
