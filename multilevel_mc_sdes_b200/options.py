"""Host-side mirror of the reference's Option classes for the hot path (drop-in surface).

Same class names, constructor kwargs, attributes and method contracts as
/root/reference/MLMC_RSCAM.py (`:NNN` = line numbers there):

    Option :777-1042, GBM_Option :1317-1363, Merton_Option :1248-1315,
    Euro_/Asian_/Lookback_/Digital_GBM :1367-1481, *_Merton :1636-1704

What differs: `looper` / `payoff` run on the GPU through libmlmcb.so (include/mlmcb.h) instead of
numpy loops; randomness is Philox addressed by (seed, level, global path id) instead of numpy's
global MT19937, so each Option carries a seed and a per-level path-offset counter (every looper
call consumes fresh ids).  `mlmc` is the reference driver restated line by line
(host code; it only decides how many samples each level gets), None-safe for alpha_0/beta_0.

Extra constructor kwargs (keyword-only, all optional): seed, device, sampler ("f64bm" - the default: 52-bit
uniforms and an fp64 Box-Muller, i.e. the precision of the reference's numpy normals - or "f32bm": the opt-in fast
sampler, 32-bit uniforms and fp32 Box-Muller normals promoted to fp64, about twice the throughput),
fp32 (path arithmetic in fp32, opt-in), scheme ("em" | "milstein": the time-stepper the demo notebook
swaps in through `opt.sde`, ExampleUsages.ipynb cells 1 and 6), distributed (shard each level call over the
torch.distributed world and all-reduce the 7 sums).
"""
import ctypes as C
import math
import sys

import numpy as np

from . import _lib
from ._lib import GBM, MERTON, EURO, ASIAN, LOOKBACK, DIGITAL, AMERICAN

_SAMPLERS = {"f32bm": _lib.SAMPLER_F32BM, "f64bm": _lib.SAMPLER_F64BM}


try:  # what scipy.stats.norm.cdf evaluates (the reference's BS() uses norm.cdf, :35)
    from scipy.special import ndtr as _norm_cdf
except Exception:  # pragma: no cover - scipy is present in this image
    def _norm_cdf(x):
        return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _stream_handle(device):
    """cudaStream_t of torch's current stream if torch is loaded, else the default stream."""
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_stream(device).cuda_stream)
    return 0


# ------------------------------------------------------------------------------------------------
# MLMC driver, restated from Option.mlmc :960-1042.  `sweep(requests, M)` must return, for each
# (Nl, l) in `requests`, the 7-vector Option.looper would return; everything else is host numpy
# exactly as in the reference (same float N, same extrapolation fix, same regressions, same
# convergence test).
# ------------------------------------------------------------------------------------------------
def mlmc_driver(sweep, eps, M, N0, alpha_0, beta_0):
    alpha = max(0, alpha_0) if alpha_0 is not None else 0       # :982 (None-safe)
    beta = max(0, beta_0) if beta_0 is not None else 0          # :983
    L = 2
    V = np.zeros(L + 1)
    N = np.zeros(L + 1)
    dN = N0 * np.ones(L + 1)
    sums = np.zeros((7, L + 1))
    sqrt_h = np.sqrt(M ** (np.arange(0, L + 1)))

    while np.sum(dN) > 0:                                       # :993
        requests = [(int(dN[l]), l) for l in range(L + 1) if dN[l] > 0]      # :994-997
        for (num, l), s in zip(requests, sweep(requests, M)):
            sums[:, l] += s

        N += dN                                                 # :999
        Yl = np.abs(sums[0, :]) / N
        V = np.maximum((sums[1, :] / N) - (Yl) ** 2, 0)

        Yl[3:] = np.maximum(Yl[3:], Yl[2:L] * M ** (-alpha))    # :1004
        V[3:] = np.maximum(V[3:], V[2:L] * M ** (-beta))        # :1005

        if alpha_0 is None:                                     # :1007-1013
            X = np.ones((L, 2))
            X[:, 0] = np.arange(1, L + 1)
            with np.errstate(divide="ignore"):
                a = np.linalg.lstsq(X, np.log(Yl[1:]), rcond=None)[0]
            alpha = max(0.5, -a[0] / np.log(M))
        if beta_0 is None:                                      # :1014-1018
            X = np.ones((L, 2))
            X[:, 0] = np.arange(1, L + 1)
            with np.errstate(divide="ignore"):
                b = np.linalg.lstsq(X, np.log(V[1:]), rcond=None)[0]
            beta = max(0.5, -b[0] / np.log(M))

        sqrt_V = np.sqrt(V)
        Nl_new = np.ceil((2 * eps ** -2) * np.sum(sqrt_V * sqrt_h) * (sqrt_V / sqrt_h))   # :1021
        dN = np.maximum(0, Nl_new - N)

        if sum(dN > 0.01 * N) == 0:                             # :1024
            if max(Yl[-2] / (M ** alpha), Yl[-1]) > (M ** alpha - 1) * eps * np.sqrt(0.5):   # :1025
                L += 1
                V = np.concatenate((V, np.zeros(1)), axis=0)
                N = np.concatenate((N, N0 * np.zeros(1)), axis=0)
                dN = np.concatenate((dN, N0 * np.ones(1)), axis=0)
                sqrt_h = np.concatenate((sqrt_h, [np.sqrt(M ** L)]), axis=0)
                sums = np.concatenate((sums, np.zeros((7, 1))), axis=1)
    return sums, N, alpha, beta


class Option:
    """Base class (reference Option :777-1042): holds alpha_0, beta_0, X0, K, T, r."""

    _model = None
    _payoff_kind = None
    _has_anti = False             # Euro_GBM / Asian_GBM bind anti_payoff + anti_path (:1379-1381, :1406-1408)

    def __init__(self, alpha_0=None, beta_0=None, X0=100, K=100, T=1, r=0.05, *,
                 seed=0, device=0, sampler="f64bm", fp32=False, scheme="em", distributed=False, verbose=True):
        self.alpha_0 = alpha_0
        self.beta_0 = beta_0
        self.X0 = X0
        self.r = r
        self.K = K
        self.T = T
        if sampler not in _SAMPLERS:
            raise ValueError(f"sampler must be one of {sorted(_SAMPLERS)}")
        self.seed = int(seed)
        self.device = int(device)
        self.sampler = sampler
        self.fp32 = bool(fp32)
        if scheme not in ("em", "milstein"):
            raise ValueError('scheme must be "em" or "milstein"')
        self.scheme = scheme
        self.distributed = bool(distributed)
        self.verbose = verbose
        self._next_path = {}          # level -> next unused global path id

    # -- virtuals the reference declares (:816-915) ------------------------------------------
    def anti_payoff(self, N_loop, l, M):
        """(0.5*(Pf+Pa), Pc) of the antithetic estimator (anti_EA_payoff :292-317); products without
        one raise NotImplementedError as in the reference (:844)."""
        if not self._has_anti:
            raise NotImplementedError("Option instance has no implemented anti_payoff method")
        return self._payoff_arrays(N_loop, l, M, _lib.ANTITHETIC)

    def path(self, N_loop, l, M):
        """What the bound reference path function returns for N_loop fresh coupled paths:
        (Xf,Xc) final values :349, (Af/T,Ac/T) :386 for Asian, (Xf,Mf,Xc,Mc) :426 for lookbacks,
        (Xf,Xc) each (N,3) for the American put :1615."""
        if self._model is None or self._payoff_kind is None:
            raise NotImplementedError("Option instance has no implemented path method")        # :859
        raw = self._raw_paths(N_loop, l, M, 0)
        if self._payoff_kind == AMERICAN:
            return raw[0:3].T.copy(), raw[3:6].T.copy()
        if self._payoff_kind == LOOKBACK:
            return raw[0], raw[1], raw[2], raw[3]
        return raw[0], raw[2]

    def anti_path(self, N_loop, l, M):
        """(Xf,Xa,Xc) :217 or (Af/T,Aa/T,Ac/T) :290; at l==0 the antithetic item equals the fine one."""
        if not self._has_anti:
            raise NotImplementedError("Option instance has no implemented anti_path method")    # :874
        raw = self._raw_paths(N_loop, l, M, _lib.ANTITHETIC)
        return raw[0], raw[1], raw[2]

    def _raw_paths(self, N_loop, l, M, extra_flags):
        import torch
        n = int(N_loop)
        start = self._take_ids(l, n)
        flags = self._flags() | extra_flags
        rows = _lib.load().mlmcb_path_rows(self._payoff_kind, flags)
        dev = torch.device("cuda", self.device)
        raw = torch.empty((rows, n), dtype=torch.float64, device=dev)
        out7 = torch.empty(7, dtype=torch.float64, device=dev)
        p = self._params((int(l),))
        _lib.check(_lib.load().mlmcb_level_paths(
            C.byref(p), self._model, self._payoff_kind, int(l), int(M), n, self.seed, start, flags,
            self.device, _stream_handle(self.device), out7.data_ptr(), raw.data_ptr()))
        return raw.cpu().numpy()

    def sde(self, X, dt, t_, dW, dJ=0):
        raise NotImplementedError("Option instance has no implemented sde method")

    def BS(self):
        raise NotImplementedError("Option instance has no implemented BS method")

    def asset_plot(self, L=6, M=2):
        raise NotImplementedError("Option instance has no implemented asset_plot method")

    # -- plumbing ----------------------------------------------------------------------------
    def _params(self, levels=()):
        """struct mlmcb_params for a call that touches `levels` (only Amer_GBM looks at them)."""
        p = _lib.Params()
        p.X0 = float(self.X0)
        p.K = float(self.K) if self.K is not None else 0.0
        p.T = float(self.T)
        p.r = float(self.r)
        p.sig = float(self.sig)
        if self._model == GBM:      # the kernel forms mu = r + drift (:1347); keep it bit-identical
            d = getattr(self, "_drift", 0.0)
            p.drift = float(d if self.r + d == self.mu else self.mu - self.r)
        else:
            p.drift = 0.0
        p.lam = float(getattr(self, "lam", 0.0))
        p.jumpmean = float(getattr(self, "jumpmean", 0.0))
        p.jumpstd = float(getattr(self, "jumpstd", 0.0))
        p.beta = float(getattr(self, "beta", 0.5826))
        p.disc = float(np.exp(-self.r * self.T))                # numpy's value, as the reference :80
        p.J_bar = float(getattr(self, "J_bar", 0.0))
        p.boundary_c = float(getattr(self, "boundary_c", 0.0))
        return p

    def _flags(self):
        return (_SAMPLERS[self.sampler] | (_lib.FP32_PATHS if self.fp32 else 0)
                | (_lib.MILSTEIN if self.scheme == "milstein" else 0))

    def _check_product(self):
        if self._model is None or self._payoff_kind is None:
            raise NotImplementedError("Option instance has no implemented payoff method")   # :829

    def _take_ids(self, l, n):
        start = self._next_path.get(l, 0)
        self._next_path[l] = start + int(n)
        return start

    def _shard(self, n, start):
        """Contiguous slice of [start, start+n) for this rank."""
        if not self.distributed:
            return n, start
        import torch.distributed as dist
        g, G = dist.get_rank(), dist.get_world_size()
        lo, hi = (g * n) // G, ((g + 1) * n) // G
        return hi - lo, start + lo

    def _nccl(self):
        """True when the all-reduce of the sums runs on the device (torch.distributed over NCCL)."""
        if not self.distributed:
            return False
        import torch.distributed as dist
        return dist.get_backend() == "nccl"

    def _allreduce(self, rows):
        """Sum host rows (k,7) over the world (gloo: the CPU test path; NCCL sweeps never leave the device,
        see _sweep)."""
        if not self.distributed:
            return rows
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(np.ascontiguousarray(rows), dtype=torch.float64)
        if dist.get_backend() == "nccl":
            t = t.to(f"cuda:{self.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    # -- the per-path function (reference payoff :62-136) ------------------------------------
    def payoff(self, N_loop, l, M):
        """(Pf, Pc) for N_loop fresh coupled paths at level l; at l==0 the second item is the raw
        coarse path output, as in the reference (:82)."""
        self._check_product()
        return self._payoff_arrays(N_loop, l, M, 0)

    def _payoff_arrays(self, N_loop, l, M, extra_flags):
        import torch
        n = int(N_loop)
        start = self._take_ids(l, n)
        dev = torch.device("cuda", self.device)
        Pf = torch.empty(n, dtype=torch.float64, device=dev)
        Pc = torch.empty(n, dtype=torch.float64, device=dev)
        out7 = torch.empty(7, dtype=torch.float64, device=dev)
        p = self._params((int(l),))
        _lib.check(_lib.load().mlmcb_level_sums(
            C.byref(p), self._model, self._payoff_kind, int(l), int(M), n, self.seed, start,
            self._flags() | extra_flags,
            self.device, _stream_handle(self.device), out7.data_ptr(), Pf.data_ptr(), Pc.data_ptr()))
        return Pf.cpu().numpy(), Pc.cpu().numpy()

    # -- the level function (reference looper :919-957) --------------------------------------
    def looper(self, Nl, l, M, anti, Npl=10 ** 4):
        """suml (7,) float64 = [sum dP, sum dP^2, sum Pf, sum Pf^2, sum Pc, sum Pc^2, sum Pc*Pf];
        [sum Pf, sum Pf^2, sum Pf, sum Pf^2, 0,0,0] at l==0.  One kernel launch; Npl is accepted
        for signature compatibility (the GPU needs no chunking)."""
        return self._sweep([(int(Nl), int(l))], M, anti=(anti == True))[0]  # noqa: E712 (reference spelling :941)

    def looper_async(self, Nl, l, M, anti=False, out=None):
        """Device-side variant of `looper`: launches on torch's current stream and returns the seven sums
        as a CUDA float64 tensor without synchronising (no D2H copy).  With `distributed=True` this rank
        simulates its slice only and the caller owns the all-reduce."""
        self._check_product()
        if anti and not self._has_anti:
            raise NotImplementedError("Option instance has no implemented anti_payoff method")
        import torch
        n = int(Nl)
        start = self._take_ids(l, n)
        n_loc, start_loc = self._shard(n, start)
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty(7, dtype=torch.float64, device=dev)
        p = self._params((int(l),))
        _lib.check(_lib.load().mlmcb_level_sums(
            C.byref(p), self._model, self._payoff_kind, int(l), int(M), n_loc, self.seed, start_loc,
            self._flags() | (_lib.ANTITHETIC if anti else 0), self.device, _stream_handle(self.device),
            out.data_ptr(), None, None))
        return out

    def _sweep(self, requests, M, anti=False):
        """requests: [(Nl, l), ...] -> list of 7-vectors.  The whole sweep is one kernel launch (per-level
        block ranges); one D2H copy brings every row back."""
        self._check_product()
        if anti and not self._has_anti:
            raise NotImplementedError("Option instance has no implemented anti_payoff method")   # :844
        k = len(requests)
        if k == 0:
            return []
        levels = (C.c_int * k)(*[l for _, l in requests])
        counts, offsets = [], []
        for n, l in requests:
            start = self._take_ids(l, n)
            n_loc, start_loc = self._shard(n, start)
            counts.append(n_loc)
            offsets.append(start_loc)
        n_arr = (C.c_uint64 * k)(*counts)
        o_arr = (C.c_uint64 * k)(*offsets)
        p = self._params(tuple(int(l) for _, l in requests))
        flags = self._flags() | (_lib.ANTITHETIC if anti else 0)
        if self._nccl():
            # multi-GPU: the kernel writes its rows into the buffer NCCL reduces in place; one D2H after it
            import torch
            import torch.distributed as dist
            dev = torch.device("cuda", self.device)
            buf = torch.empty((k, 8), dtype=torch.float64, device=dev)
            _lib.check(_lib.load().mlmcb_level_sweep(
                C.byref(p), self._model, self._payoff_kind, int(M), k, levels, n_arr, o_arr, self.seed, flags,
                self.device, _stream_handle(self.device), buf.data_ptr()))
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            out = buf.cpu().numpy()[:, :7]
            return [out[i].copy() for i in range(k)]
        out = np.zeros((k, 7))
        _lib.check(_lib.load().mlmcb_level_sweep_host(
            C.byref(p), self._model, self._payoff_kind, int(M), k, levels, n_arr, o_arr, self.seed, flags,
            self.device, _stream_handle(self.device), out.ctypes.data_as(C.POINTER(C.c_double))))
        out = self._allreduce(out)
        return [out[i].copy() for i in range(k)]

    # -- replay (parity check (a)) -----------------------------------------------------------
    def payoff_replay(self, draws, l, M, arith="fast", anti=False):
        """Feed the reference's own draws.  GBM: `draws` = Z (M**l, N) array, row j the
        np.random.randn(N) of step j+1.  Merton: dict with zW/offW, zQ/offQ, E/offE (CSR, the
        order the reference consumes them).  Returns (Pf, Pc) per path."""
        self._check_product()
        import torch
        dev = torch.device("cuda", self.device)
        ar = {"fast": _lib.ARITH_FAST, "exact": _lib.ARITH_EXACT}[arith]
        if self.scheme == "milstein":
            ar |= _lib.MILSTEIN
        if anti:
            if not self._has_anti:
                raise NotImplementedError("Option instance has no implemented anti_payoff method")
            ar |= _lib.ANTITHETIC
        p = self._params((int(l),))
        st = _stream_handle(self.device)
        if self._model == GBM:
            Z = torch.as_tensor(np.ascontiguousarray(draws, dtype=np.float64), device=dev)
            if Z.shape[0] != int(M) ** int(l):
                raise ValueError("Z must have M**l rows")
            n = Z.shape[1]
            Pf = torch.empty(n, dtype=torch.float64, device=dev)
            Pc = torch.empty(n, dtype=torch.float64, device=dev)
            _lib.check(_lib.load().mlmcb_replay_gbm(C.byref(p), self._payoff_kind, int(l), int(M), n, Z.data_ptr(),
                                                    ar, self.device, st, Pf.data_ptr(), Pc.data_ptr(), None))
        else:
            t = {k: torch.as_tensor(np.ascontiguousarray(draws[k]), device=dev)
                 for k in ("zW", "offW", "zQ", "offQ", "E", "offE")}
            # empty streams still need a valid pointer
            for k in ("zW", "zQ", "E"):
                if t[k].numel() == 0:
                    t[k] = torch.zeros(1, dtype=torch.float64, device=dev)
            n = t["offW"].numel() - 1
            Pf = torch.empty(n, dtype=torch.float64, device=dev)
            Pc = torch.empty(n, dtype=torch.float64, device=dev)
            _lib.check(_lib.load().mlmcb_replay_merton(
                C.byref(p), self._payoff_kind, int(l), int(M), n, t["zW"].data_ptr(), t["offW"].data_ptr(),
                t["zQ"].data_ptr(), t["offQ"].data_ptr(), t["E"].data_ptr(), t["offE"].data_ptr(),
                ar, self.device, st, Pf.data_ptr(), Pc.data_ptr(), None))
        return Pf.cpu().numpy(), Pc.cpu().numpy()

    # -- MLMC driver (reference mlmc :960-1042) ----------------------------------------------
    def mlmc(self, eps, M=2, anti=False, N0=10 ** 3, warm_start=True):
        """Returns (sums (7,L+1) float64, N (L+1,) float64); price = sum(sums[0,:]/N)."""
        use_anti = anti == True  # noqa: E712
        if use_anti and not self._has_anti:
            raise NotImplementedError("Option instance has no implemented anti_payoff method")
        sums, N, alpha, beta = mlmc_driver(lambda reqs, M_: self._sweep(reqs, M_, anti=use_anti), eps, M, N0,
                                           self.alpha_0, self.beta_0)
        if self.verbose:
            print(f'Estimated alpha = {alpha}')                 # :1034
            print(f'Estimated beta = {beta}')
        if warm_start:                                          # :1037-1041
            self.alpha_0 = alpha
            self.beta_0 = beta
            if self.verbose:
                print(f'    Saved estimated alpha_0 = {alpha}')
                print(f'    Saved estimated beta_0 = {beta}')
        return sums, N

    # -- numeric core of Giles_plot's fixed-level sweep (:1775-1777) -------------------------
    def level_sweep(self, Lmax=8, Nsamples=10 ** 6, M=2, anti=False):
        """sums (7, Lmax+1) of looper(Nsamples, l, M, anti) for l = 0..Lmax, issued as one sweep."""
        rows = self._sweep([(int(Nsamples), l) for l in range(Lmax + 1)], M, anti=anti)
        return np.stack(rows, axis=1)


class GBM_Option(Option):
    """dS = mu*S*dt + sig*S*dW, mu = drift + r  (reference :1317-1363)."""
    _model = GBM

    def __init__(self, drift=0, sig=0.2, **kwargs):
        super().__init__(**kwargs)
        self.sig = sig
        r = self.r
        self._drift = drift
        self.mu = (r + drift)                                   # :1347

    def sde(self, X, dt, t_, dW):
        if self.scheme == "milstein":                           # notebook cell 1, milstein_GBM
            return self.mu * X * dt + self.sig * X * dW + 0.5 * X * (self.sig ** 2) * (dW ** 2 - dt)
        return self.mu * X * dt + self.sig * X * dW             # :1363


class Merton_Option(Option):
    """Merton jump-diffusion (reference :1248-1315)."""
    _model = MERTON

    def __init__(self, lam=1, jumpmean=0.1, jumpstd=0.2, sig=0.2, **kwargs):
        super().__init__(**kwargs)
        self.lam = lam
        self.J_bar = np.exp(jumpmean + 0.5 * jumpstd ** 2) - 1   # :1291
        self.jumpstd = jumpstd
        self.jumpmean = jumpmean
        # kept for attribute parity with the reference (:1294-1295); the kernels draw their own
        # exponential gaps and jump sizes from Philox and never call these
        self.jumptime = np.random.exponential
        self.jumpsize = lambda: jumpmean + jumpstd * np.random.standard_normal()
        self.sig = sig

    def sde(self, X, dt, t_, dW, dJ=0):
        dx_ = (self.r - self.lam * self.J_bar) * X * dt + self.sig * X * dW     # :1314
        if self.scheme == "milstein":                           # notebook cell 1, milstein_Merton
            dx_ = dx_ + 0.5 * X * (self.sig ** 2) * (dW ** 2 - dt)
        return dx_ + (dx_ + X) * dJ


class _HostOnlyOption(Option):
    """Reference classes whose SDE / path / payoff are arbitrary user Python callables
    (JumpDiffusion_Option :1044-1120, Diffusion_Option :1122-1209, MyOption :1211-1244).  A Python
    lambda cannot run inside a CUDA kernel, so these stay with the reference's host path; the names
    exist here only to fail with a clear message instead of an AttributeError."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} takes user-supplied Python callables and is outside the GPU hot path; "
            "use the reference implementation for it (see DESIGN.md, out of scope)")


class JumpDiffusion_Option(_HostOnlyOption):
    pass


class Diffusion_Option(_HostOnlyOption):
    pass


class MyOption(_HostOnlyOption):
    pass


class Euro_GBM(GBM_Option):
    _payoff_kind = EURO
    _has_anti = True

    def BS(self):                                               # :1383-1393
        D1 = (np.log(self.X0 / self.K) + (self.r + 0.5 * self.sig ** 2) * self.T) / (self.sig * np.sqrt(self.T))
        D2 = D1 - self.sig * np.sqrt(self.T)
        return self.X0 * _norm_cdf(D1) - self.K * np.exp(-self.r * self.T) * _norm_cdf(D2)


class Asian_GBM(GBM_Option):
    _payoff_kind = ASIAN
    _has_anti = True

    def BS(self):                                               # Turnbull-Wakeman approximation :1411-1429
        X0, r, T, sig, K = self.X0, self.r, self.T, self.sig, self.K
        M1 = X0 * (np.exp(r * T) - 1) / (r * T)
        M2 = (2 * X0 ** 2) * (np.exp((2 * r + sig ** 2) * T) / ((r + sig ** 2) * (2 * r + sig ** 2) * T ** 2)
                              + (1 / (2 * r + sig ** 2) - np.exp(r * T) / (r + sig ** 2)) / (r * T ** 2))
        sig_A = np.sqrt(np.log(M2 / (M1 ** 2)))
        d1 = (np.log(M1 / K) + 0.5 * sig_A ** 2) / sig_A
        d2 = d1 - sig_A
        return np.exp(-r * T) * (M1 * _norm_cdf(d1) - K * _norm_cdf(d2))


class Lookback_GBM(GBM_Option):
    _payoff_kind = LOOKBACK
    beta = 0.5826                                               # :1443

    def BS(self):                                               # :1446-1456
        D1 = (self.r + 0.5 * self.sig ** 2) * self.T / (self.sig * np.sqrt(self.T))
        D2 = D1 - self.sig * np.sqrt(self.T)
        k = 0.5 * self.sig ** 2 / self.r
        return self.X0 * (_norm_cdf(D1) - _norm_cdf(-D1) * k
                          - np.exp(-self.r * self.T) * (_norm_cdf(D2) - _norm_cdf(D2) * k))


class Digital_GBM(GBM_Option):
    _payoff_kind = DIGITAL

    def BS(self):                                               # :1473-1481
        D2 = ((np.log(self.X0 / self.K) + (self.r + 0.5 * self.sig ** 2) * self.T) / (self.sig * np.sqrt(self.T))
              - self.sig * np.sqrt(self.T))
        return self.K * np.exp(-self.r * self.T) * _norm_cdf(D2)


def sqrt_boundary(c=0.8):
    """The exercise-boundary family of the demo notebook (ExampleUsages.ipynb cell 1, `exb`):
    b(self, t) = max(K*(1 - c*sqrt(T - t)), 0).  The returned callable has the reference's
    `exerciseBoundary(self, t)` signature and carries `c` so Amer_GBM can run it on the GPU."""
    def boundary(self, t):
        return np.maximum(self.K * (1 - c * np.sqrt(self.T - t)), 0)
    boundary.mlmcb_c = float(c)
    return boundary


class Amer_GBM(GBM_Option):
    """American put under GBM (reference Amer_GBM :1483-1632, Amer_payoff :138-158): barrier-crossing
    probabilities between grid points, M must be 2 (:1524).  `exerciseBoundary` is what the reference takes -
    any callable `b(self, t)` (:1495-1497, called as `exerciseBoundary(self, t)` :1561) - or `sqrt_boundary(c)` /
    the number c for the demo notebook's family K(1 - c sqrt(T-t)), which the library tabulates on the device.
    A general callable is evaluated HERE, on the host, once per level, at exactly the times the reference
    evaluates it (grid points :1561/:1568, fine mid-points :1565, coarse mid-points :1593), and only the values
    travel to the library (mlmcb_params.amer_boundary); the path simulation stays on the GPU."""
    _payoff_kind = AMERICAN

    def __init__(self, exerciseBoundary=0.8, **kwargs):
        super().__init__(**kwargs)
        self.exerciseBoundary = exerciseBoundary

    @property
    def exerciseBoundary(self):
        return self._exercise_boundary

    @exerciseBoundary.setter
    def exerciseBoundary(self, b):
        """The reference lets a user rebind `opt.exerciseBoundary` at any time: every assignment re-derives how the
        boundary reaches the kernel and drops the tabulated values of the previous one."""
        c = getattr(b, "mlmcb_c", b)
        self._boundary_tables = {}
        if callable(c):
            self.boundary_c = 0.0
            self._exercise_boundary = b
            self._tabulated = True
        else:
            self.boundary_c = float(c)
            self._exercise_boundary = sqrt_boundary(self.boundary_c)
            self._tabulated = False

    def _boundary_values(self, l):
        """b on the grid | fine mid-points | coarse mid-points of level l (M = 2), as float64, cached per level
        and parameter set."""
        key = (int(l), float(self.T), float(self.K), float(self.X0), float(self.r), float(self.sig))
        tab = self._boundary_tables.get(key)
        if tab is None:
            n, M = 2 ** int(l), 2
            dt = self.T / n                                              # :1529
            b = self.exerciseBoundary
            grid = [b(self, j * dt) for j in range(n + 1)]               # t_ = (j-1)*dt :1546, j*dt :1568
            fmid = [b(self, (j - 1) * dt + dt / 2) for j in range(1, n + 1)]                 # :1562
            cmid = [b(self, (j - M) * dt + M * dt / 2) for j in range(2, n + 1, 2)]          # :1580, :1590
            tab = np.ascontiguousarray(np.array(grid + fmid + cmid, dtype=np.float64).reshape(-1))
            if tab.size != 2 * n + 1 + n // 2:
                raise ValueError("exerciseBoundary(self, t) must return one number per call")
            self._boundary_tables[key] = tab
        return tab

    def _params(self, levels=()):
        p = super()._params(levels)
        if self._tabulated and levels:
            top = max(levels)
            ptrs = (C.c_void_p * (top + 1))()
            keep = []
            for l in set(levels):
                tab = self._boundary_values(l)
                keep.append(tab)
                ptrs[l] = tab.ctypes.data
            p.amer_boundary = C.cast(ptrs, C.POINTER(C.c_void_p))
            p.amer_levels = top + 1
            p._keep_alive = (ptrs, keep)                                 # host memory the library reads during the call
        return p


class Euro_Merton(Merton_Option):
    _payoff_kind = EURO

    def BS(self, n=100):                                        # :1648-1664
        k = np.arange(0, n)
        sig_n = np.sqrt(self.sig ** 2 + (self.jumpstd ** 2) * k / self.T)
        r_n = self.r - self.lam * self.J_bar + k * np.log((1 + self.J_bar) / self.T)
        D1 = (np.log(self.X0 / self.K) + (r_n + 0.5 * sig_n ** 2) * self.T) / (sig_n * np.sqrt(self.T))
        D2 = D1 - sig_n * np.sqrt(self.T)
        bs = self.X0 * _norm_cdf(D1) - self.K * np.exp(-r_n * self.T) * _norm_cdf(D2)
        discount = np.exp(-self.lam * (1 + self.J_bar) * self.T)
        fact = np.array([math.factorial(int(i)) for i in k], dtype=np.float64)
        return discount * np.sum(((self.lam * (1 + self.J_bar) * self.T) ** k / fact) * bs)


class Asian_Merton(Merton_Option):
    _payoff_kind = ASIAN


class Lookback_Merton(Merton_Option):
    _payoff_kind = LOOKBACK
    beta = 0.5826                                               # :1691


class Digital_Merton(Merton_Option):
    _payoff_kind = DIGITAL
