"""ctypes binding of libmlmcb.so (include/mlmcb.h).  The library is the product's only compute
path: if it is missing or fails to load this module raises - there is no Python/CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLMCB_LIB") or os.path.join(_HERE, "csrc", "libmlmcb.so")   # MLMCB_LIB: kernel experiments

GBM, MERTON = 0, 1
EURO, ASIAN, LOOKBACK, DIGITAL, AMERICAN = 0, 1, 2, 3, 4
SAMPLER_F64BM, SAMPLER_F32BM = 0, 1      # flags == 0: the fp64 sampler; fp32 is the opt-in
FP32_PATHS = 0x10
ANTITHETIC = 0x20
MILSTEIN = 0x40
ARITH_FAST, ARITH_EXACT = 0, 1
OK, EINVAL, ECUDA, ENOTIMPL = 0, -1, -2, -3
ABI_VERSION = 102


class Params(C.Structure):
    """struct mlmcb_params (include/mlmcb.h)."""
    _fields_ = [(k, C.c_double) for k in
                ("X0", "K", "T", "r", "sig", "drift", "lam", "jumpmean", "jumpstd", "beta", "disc", "J_bar",
                 "boundary_c")] + [("amer_boundary", C.POINTER(C.c_void_p)), ("amer_levels", C.c_int),
                                   ("reserved_", C.c_int)]


class MlmcbError(RuntimeError):
    pass


_dp = C.POINTER(C.c_double)
_u64 = C.c_uint64
_vp = C.c_void_p

# name -> (restype, argtypes); also the list tests check against the header
SIGNATURES = {
    "mlmcb_level_sums": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _u64, _u64, _u64,
                                   C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mlmcb_path_rows": (C.c_int, [C.c_int, C.c_int]),
    "mlmcb_level_paths": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _u64, _u64, _u64,
                                    C.c_int, C.c_int, _vp, _vp, _vp]),
    "mlmcb_level_sums_host": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _u64, _u64, _u64,
                                        C.c_int, C.c_int, _vp, _dp]),
    "mlmcb_level_sweep_host": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(_u64), C.POINTER(_u64), _u64,
                                         C.c_int, C.c_int, _vp, _dp]),
    "mlmcb_level_sweep": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(_u64), C.POINTER(_u64), _u64,
                                    C.c_int, C.c_int, _vp, _vp]),
    "mlmcb_replay_gbm": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, _u64, _vp, C.c_int, C.c_int,
                                   _vp, _vp, _vp, _vp]),
    "mlmcb_replay_merton": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, _u64, _vp, _vp, _vp, _vp,
                                      _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mlmcb_sampler_probe": (C.c_int, [C.c_int, _u64, C.c_int, _u64, _u64, _u64, C.c_int, _vp, _vp]),
    "mlmcb_packed_probe": (C.c_int, [C.c_int, _u64, C.c_int, C.c_int, _u64, _u64, C.c_int, _vp, _vp]),
    "mlmcb_jump_probe": (C.c_int, [C.c_int, _u64, C.c_int, C.c_double, _u64, _u64, _u64, C.c_int, _vp, _vp]),
    "mlmcb_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mlmcb_pipe_probe": (C.c_int, [C.c_int, C.c_int, _dp, C.POINTER(C.c_float)]),
    "mlmcb_mix_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "mlmcb_launch_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mlmcb_params_fill_derived": (None, [C.POINTER(Params)]),
    "mlmcb_last_error": (C.c_char_p, []),
    "mlmcb_version": (C.c_int, []),
}

_lib = None


def load():
    """Load libmlmcb.so (once).  Raises MlmcbError if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MlmcbError(
            f"{LIB_PATH} is missing - build it with `python -m multilevel_mc_sdes_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.mlmcb_version() != ABI_VERSION:      # a stale build would read struct mlmcb_params with another layout
        raise MlmcbError(f"{LIB_PATH} has ABI version {lib.mlmcb_version()}, this package needs {ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    msg = load().mlmcb_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOTIMPL:
        raise NotImplementedError(msg)
    raise MlmcbError(msg)


def philox4x32_10(ctr, key):
    """Host-side Philox4x32-10 (the same round function the kernels run)."""
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    load().mlmcb_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)
