"""numpy restatement of the per-level coupled sampler of MLMC_RSCAM.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file, and only as the checker / the timed CPU baseline.  The product
package (`multilevel_mc_sdes_b200`) never imports it: its hot path is the CUDA library.

Parity status: PINNED.  `tests/test_oracle_vs_reference.py` (runs where `/root/reference`
exists) checks every function below bit-for-bit against the unmodified reference with the
same `np.random.seed`, and `tests/golden/*.npz` (made by `oracle/make_golden.py` from the
reference itself) pins it where the reference tree is absent.

All `:NNN` citations are line numbers in /root/reference/MLMC_RSCAM.py.

Randomness is abstracted behind a *draw source* so the same restatement serves three uses:
  * `GlobalDraws`   - numpy legacy global RNG, consumed in exactly the reference's order
                      (np.random.randn :343/:460, np.random.standard_normal :1295,
                      np.random.exponential :1294) -> bit-identical to the reference for a seed;
  * `RecordingDraws`- same, and logs every draw into per-path ragged streams (replay fixtures);
  * `ReplayDraws`   - feeds previously recorded draws back.
"""
import numpy as np

GBM, MERTON = 0, 1
EURO, ASIAN, LOOKBACK, DIGITAL, AMERICAN = 0, 1, 2, 3, 4
MODEL_NAMES = {GBM: "GBM", MERTON: "Merton"}
PAYOFF_NAMES = {EURO: "Euro", ASIAN: "Asian", LOOKBACK: "Lookback", DIGITAL: "Digital"}


class Params:
    """Parameter bundle; field names follow the reference attributes (Option :796-813,
    GBM_Option :1330-1347, Merton_Option :1273-1296, Lookback beta :1443/:1691)."""

    def __init__(self, X0=100, K=100, T=1, r=0.05, sig=0.2, drift=0, lam=1, jumpmean=0.1,
                 jumpstd=0.2, beta=0.5826, scheme="em", boundary_c=0.8):
        self.scheme = scheme                                        # "em" | "milstein" (notebook cell 1)
        self.boundary_c = boundary_c                                # exercise boundary K*(1-c*sqrt(T-t)) (notebook `exb`)
        self.X0, self.K, self.T, self.r, self.sig, self.drift = X0, K, T, r, sig, drift
        self.lam, self.jumpmean, self.jumpstd, self.beta = lam, jumpmean, jumpstd, beta
        self.mu = r + drift                                         # :1347
        self.J_bar = np.exp(jumpmean + 0.5 * jumpstd ** 2) - 1      # :1291


# ----------------------------------------------------------------------------- draw sources
class GlobalDraws:
    """numpy legacy global stream, same calls as the reference."""

    def normal_vec(self, n):          # np.random.randn(N_loop) :343 :374 :417
        return np.random.randn(n)

    def normal_w(self):               # np.random.randn() :460 :479
        return np.random.randn()

    def normal_q(self):               # np.random.standard_normal() inside jumpsize :1295
        return np.random.standard_normal()

    def expo(self, scale):            # jumptime = np.random.exponential :1294, called :454 :474
        return np.random.exponential(scale=scale)

    def next_path(self):
        pass


class RecordingDraws(GlobalDraws):
    """Logs the scalar draws of the jump-diffusion path functions per path."""

    def __init__(self):
        self.zw, self.zq, self.ex = [[]], [[]], [[]]

    def normal_w(self):
        v = np.random.randn()
        self.zw[-1].append(v)
        return v

    def normal_q(self):
        v = np.random.standard_normal()
        self.zq[-1].append(v)
        return v

    def expo(self, scale):
        # stored as drawn, i.e. already scaled by 1/lam (an inter-arrival time)
        v = np.random.exponential(scale=scale)
        self.ex[-1].append(v)
        return v

    def next_path(self):
        self.zw.append([])
        self.zq.append([])
        self.ex.append([])

    def packed(self):
        """-> dict of flat arrays + int64 offsets (CSR layout, one row per path)."""
        zw, zq, ex = self.zw, self.zq, self.ex
        if zw and not zw[-1] and not ex[-1]:
            zw, zq, ex = zw[:-1], zq[:-1], ex[:-1]

        def csr(rows):
            off = np.zeros(len(rows) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(r) for r in rows])
            flat = np.array([v for r in rows for v in r], dtype=np.float64)
            return flat, off

        fw, ow = csr(zw)
        fq, oq = csr(zq)
        fe, oe = csr(ex)
        return dict(zW=fw, offW=ow, zQ=fq, offQ=oq, E=fe, offE=oe)


class ReplayDraws:
    """Feeds recorded jump-diffusion draws back (E holds the *scaled* exponentials)."""

    def __init__(self, packed):
        self.p = packed
        self.path = 0
        self._reset()

    def _reset(self):
        self.iw = int(self.p["offW"][self.path])
        self.iq = int(self.p["offQ"][self.path])
        self.ie = int(self.p["offE"][self.path])

    def normal_w(self):
        v = self.p["zW"][self.iw]
        self.iw += 1
        return v

    def normal_q(self):
        v = self.p["zQ"][self.iq]
        self.iq += 1
        return v

    def expo(self, scale):
        v = self.p["E"][self.ie]
        self.ie += 1
        return v

    def next_path(self):
        self.path += 1
        if self.path < len(self.p["offW"]) - 1:
            self._reset()


# ----------------------------------------------------------------------------- SDE increments
def gbm_increment(p, X, dt, dW):
    """GBM_Option.sde :1363  (t_ is ignored by the reference); with scheme="milstein" the
    `milstein_GBM` the demo notebook binds in its place (ExampleUsages.ipynb cell 1)."""
    if p.scheme == "milstein":
        return p.mu * X * dt + p.sig * X * dW + 0.5 * X * (p.sig ** 2) * (dW ** 2 - dt)
    return p.mu * X * dt + p.sig * X * dW


def merton_increment(p, X, dt, dW, dJ=0):
    """Merton_Option.sde :1314-1315 / notebook `milstein_Merton`."""
    if p.scheme == "milstein":
        dx_ = (p.r - p.lam * p.J_bar) * X * dt + p.sig * X * dW + 0.5 * X * (p.sig ** 2) * (dW ** 2 - dt)
    else:
        dx_ = (p.r - p.lam * p.J_bar) * X * dt + p.sig * X * dW
    return dx_ + (dx_ + X) * dJ


# ----------------------------------------------------------------------------- GBM coupled paths
def gbm_paths(p, functional, l, M, Z=None, n=None):
    """diffusion_path :320-349 / diffusion_path_avg :351-386 / diffusion_path_min :388-426.

    functional: EURO/DIGITAL -> final values, ASIAN -> trapezoid averages, LOOKBACK -> (X, min).
    Z: optional (M**l, n) array of standard normals in the reference's step-major order
       (row j-1 is the `randn(N_loop)` call of step j).  If None the numpy global RNG is used.
    Returns what the reference path function returns: (Xf,Xc) | (Af/T,Ac/T) | (Xf,Mf,Xc,Mc).
    """
    T, X0 = p.T, p.X0
    Nsteps = M ** l
    dt = T / Nsteps
    sqrt_dt = np.sqrt(dt)
    if Z is not None:
        n = Z.shape[1]
    Xf = X0 * np.ones(n)
    Xc = X0 * np.ones(n)
    dWc = np.zeros(n)
    if functional == ASIAN:
        Af = 0.5 * dt * Xf                                          # :369
        Ac = 0.5 * M * dt * Xc                                      # :370
    elif functional == LOOKBACK:
        Mf = X0 * np.ones(n)                                        # :411
        Mc = X0 * np.ones(n)                                        # :412
    for j in range(1, Nsteps + 1):
        z = Z[j - 1] if Z is not None else np.random.randn(n)
        dWf = z * sqrt_dt                                           # :343
        dWc = dWc + dWf                                             # :344
        Xf += gbm_increment(p, Xf, dt, dWf)                         # :345
        if functional == ASIAN:
            Af += Xf * dt                                           # :377
        elif functional == LOOKBACK:
            Mf = np.minimum(Xf, Mf)                                 # :420
        if j % M == 0:
            Xc += gbm_increment(p, Xc, M * dt, dWc)                 # :347
            if functional == ASIAN:
                Ac += Xc * M * dt                                   # :380
            elif functional == LOOKBACK:
                Mc = np.minimum(Xc, Mc)                             # :423
            dWc = np.zeros(n)                                       # :348
    if functional == ASIAN:
        Af -= 0.5 * Xf * dt                                         # :383
        Ac -= 0.5 * Xc * M * dt                                     # :384
        return Af / T, Ac / T
    if functional == LOOKBACK:
        return Xf, Mf, Xc, Mc
    return Xf, Xc


# ----------------------------------------------------------------------------- GBM antithetic paths
def gbm_anti_paths(p, functional, l, M, Z=None, n=None):
    """diffusion_anti_path :162-217 (functional EURO) / diffusion_anti_path_avg :219-290 (ASIAN).
    Z rows follow the reference's call order: row 2k = dWf_odd of pair k, row 2k+1 = dWf_even
    (:195-196); at l==0 the single row is the :214 / :283 draw.  Returns (f, a, c)."""
    if M % 2 != 0:
        raise Exception("Cannot calculate antithetic estimator for odd coarseness factor M - please use even M")
    T, X0 = p.T, p.X0
    Nsteps = M ** l
    dt = T / Nsteps
    sqrt_dt = np.sqrt(dt)
    if Z is not None:
        n = Z.shape[1]
    draw = (lambda i: Z[i]) if Z is not None else (lambda i: np.random.randn(n))
    Xf = X0 * np.ones(n)
    Xc = X0 * np.ones(n)
    Xa = X0 * np.ones(n)
    if functional == ASIAN:
        Af = 0.5 * dt * Xf                                          # :247
        Aa = 0.5 * dt * Xa
        Ac = 0.5 * M * dt * Xc
    dWc = np.zeros(n)
    k = 0
    for j in range(2, Nsteps + 1, 2):
        dWf_odd = draw(k) * sqrt_dt                                 # :195
        dWf_even = draw(k + 1) * sqrt_dt                            # :196
        k += 2
        Xf += gbm_increment(p, Xf, dt, dWf_odd)                     # :200
        Xa += gbm_increment(p, Xa, dt, dWf_even)                    # :201
        if functional == ASIAN:
            Af += dt * Xf                                           # :263
            Aa += dt * Xa
        Xf += gbm_increment(p, Xf, dt, dWf_even)                    # :205
        Xa += gbm_increment(p, Xa, dt, dWf_odd)                     # :206
        if functional == ASIAN:
            Af += dt * Xf                                           # :271
            Aa += dt * Xa
        dWc += dWf_odd + dWf_even                                   # :208
        if j % M == 0:
            Xc += gbm_increment(p, Xc, M * dt, dWc)                 # :210
            if functional == ASIAN:
                Ac += dt * M * Xc                                   # :278
            dWc = np.zeros(n)
    if l == 0:
        Xf += gbm_increment(p, Xf, dt, draw(0) * sqrt_dt)           # :214 / :283
        if functional == ASIAN:
            Af += 0.5 * dt * Xf                                     # :284
            return Af / T, Af / T, Xc
        return Xf, Xf, Xc
    if functional == ASIAN:
        Af -= 0.5 * Xf * dt                                         # :287-289
        Aa -= 0.5 * Xa * dt
        Ac -= 0.5 * Xc * (M * dt)
        return Af / T, Aa / T, Ac / T
    return Xf, Xa, Xc


def anti_level_payoffs(p, payoff, l, M, n=None, Z=None):
    """anti_EA_payoff :292-317 for Euro_GBM / Asian_GBM -> (0.5*(Pf+Pa), Pc); l==0: (Pf, Xc)."""
    r, K, T = p.r, p.K, p.T
    Xf, Xa, Xc = gbm_anti_paths(p, functional_of(payoff), l, M, Z=Z, n=n)
    Pf = np.maximum(0, Xf - K)
    Pf = np.exp(-r * T) * Pf
    if l == 0:
        return Pf, Xc
    Pa = np.maximum(0, Xa - K)
    Pa = np.exp(-r * T) * Pa
    Pc = np.maximum(0, Xc - K)
    Pc = np.exp(-r * T) * Pc
    return 0.5 * (Pf + Pa), Pc


# ----------------------------------------------------------------------------- American put (GBM)
def exercise_boundary(p, t):
    """The demo notebook's `exb` (ExampleUsages.ipynb cell 1): max(K*(1-0.8*sqrt(T-t)), 0), with the
    0.8 exposed as p.boundary_c.  It is the callback `Amer_GBM(exerciseBoundary=...)` is given."""
    return np.maximum(p.K * (1 - p.boundary_c * np.sqrt(p.T - t)), 0)


def amer_paths(p, l, M, Z=None, n=None):
    """Amer_GBM.path :1511-1615.  Returns (Xf, Xc), each (n,3): price, probability of not having
    crossed, accumulated exercise value.  Keeps the reference's aliasing at :1577: `halfdWc=dWc`
    binds the SAME array that `dWc += dWf` (:1549) then updates in place, so at the coarse step
    halfdWc is dWc (and `halfdWc - 0.5*dWc` is 0.5*dWc)."""
    if M != 2:
        raise ValueError("M must be equal to 2 for American Option pricer. Please use M=2.")     # :1525
    T, X0, sig, K, r = p.T, p.X0, p.sig, p.K, p.r
    eb = lambda t: exercise_boundary(p, t)
    Nsteps = M ** l
    dt = T / Nsteps
    sqrt_dt = np.sqrt(dt)
    if Z is not None:
        n = Z.shape[1]
    dWc = np.zeros(n)
    Xf = np.zeros((n, 3))
    Xf[:, 0] = X0 * np.ones(n)
    Xf[:, 1] = np.ones(n)
    Xf[:, 2] = np.zeros(n)
    Xc = Xf.copy()
    halfdWc = dWc
    for j in range(1, Nsteps + 1):
        t_ = (j - 1) * dt
        z = Z[j - 1] if Z is not None else np.random.randn(n)
        dWf = sqrt_dt * z                                           # :1548
        dWc += dWf                                                  # :1549 (in place)
        Xleft = Xf[:, 0]
        Xright = Xleft + gbm_increment(p, Xleft, dt, dWf)           # :1554
        leftBarrier = eb(t_)
        tfMid = t_ + dt / 2
        rightBarrier = eb(j * dt)
        midBarrier = eb(tfMid)
        Prob1 = np.exp(-2 * np.maximum(Xleft - leftBarrier, 0) * np.maximum(Xright - rightBarrier, 0)
                       / ((sig ** 2) * dt * Xleft ** 2))             # :1566
        Xf[:, 2] += Xf[:, 1] * Prob1 * np.maximum(K - midBarrier, 0) * np.exp(-r * (tfMid))    # :1569
        Xf[:, 1] *= (1.0 - Prob1)                                   # :1571
        Xf[:, 0] = Xright                                           # :1573
        if j % M == M / 2:
            halfdWc = dWc                                           # :1577 (alias, see docstring)
        if j % M == 0:
            t_ = (j - M) * dt
            Xleft = Xc[:, 0]
            Xright = Xleft + gbm_increment(p, Xleft, M * dt, dWc)   # :1584
            Xmid = Xleft + 0.5 * (Xright - Xleft) + sig * Xleft * (halfdWc - 0.5 * dWc)     # :1585
            leftBarrier = eb(t_)
            tcMid = t_ + M * dt / 2
            rightBarrier = eb(j * dt)
            midBarrier2 = eb(tcMid)
            midBarrier1 = 0.5 * (rightBarrier + leftBarrier)
            Prob11 = np.exp(-2 * np.maximum(Xleft - leftBarrier, 0) * np.maximum(Xmid - midBarrier1, 0)
                            / (dt * (sig ** 2) * (Xleft ** 2)))     # :1600
            Prob12 = np.exp(-2 * np.maximum(Xmid - midBarrier1, 0) * np.maximum(Xright - rightBarrier, 0)
                            / (dt * (sig ** 2) * (Xleft ** 2)))     # :1601
            Prob1 = (1.0 - (1.0 - Prob11) * (1.0 - Prob12))         # :1603
            Xc[:, 2] += Xc[:, 1] * Prob1 * np.maximum(K - midBarrier2, 0) * np.exp(-r * tcMid)   # :1605
            Xc[:, 1] *= (1 - Prob1)                                 # :1607
            Xc[:, 0] = Xright
            dWc = np.zeros(n)                                       # :1612
    return Xf, Xc


def amer_level_payoffs(p, l, M, n=None, Z=None):
    """Amer_payoff :138-158 over Amer_GBM.path; at l==0 the second item is 0 (:155)."""
    K, T, r = p.K, p.T, p.r
    Xf, Xc = amer_paths(p, l, M, Z=Z, n=n)
    Pf = Xf[:, 2] + Xf[:, 1] * np.maximum(K - Xf[:, 0], 0) * np.exp(-r * T)
    if l == 0:
        return Pf, 0
    Pc = Xc[:, 2] + Xc[:, 1] * np.maximum(K - Xc[:, 0], 0) * np.exp(-r * T)
    return Pf, Pc


# ----------------------------------------------------------------------------- Merton coupled paths
def merton_paths(p, functional, l, M, n, draws=None):
    """JD_path :428-492 / JD_path_avg :494-571 / JD_path_min :573-651 (jump-adapted EM).

    Scalar loop per path, exactly the reference's control flow and draw order:
    E, then per fine step { (N_w, N_q, E)*  N_w }.
    """
    d = draws if draws is not None else GlobalDraws()
    lam, X0, T = p.lam, p.X0, p.T
    Nsteps = M ** l
    out_f = np.zeros(n)
    out_c = np.zeros(n)
    min_f = np.zeros(n)
    min_c = np.zeros(n)
    for num in range(n):
        dWc = 0
        tau = 0
        tf = 0
        tc = 0
        dtc = 0
        Sf = X0
        Sc = X0
        avg_f = 0
        avg_c = 0
        mf = X0
        mc = X0
        tau += d.expo(1 / lam)                                      # :454
        for j in range(1, Nsteps + 1):
            tn = j * T / Nsteps                                     # :456
            while tau < tn:                                         # :457
                dt = tau - tf
                dtc += dt
                dWf = d.normal_w() * np.sqrt(dt)                    # :460
                dWc += dWf
                dJ = np.exp(p.jumpmean + p.jumpstd * d.normal_q()) - 1   # :462 with :1295
                if functional == ASIAN:                             # :532-540
                    avg_f += 0.5 * dt * Sf
                    Sf += merton_increment(p, Sf, dt, dWf)
                    avg_f += 0.5 * dt * Sf
                    Sf += Sf * dJ
                    avg_c += 0.5 * dtc * Sc
                    Sc += merton_increment(p, Sc, dtc, dWc)
                    avg_c += 0.5 * dtc * Sc
                    Sc += Sc * dJ
                else:                                               # :465-468 / :615-622
                    Sf += merton_increment(p, Sf, dt, dWf, dJ)
                    Sc += merton_increment(p, Sc, dtc, dWc, dJ)
                    if functional == LOOKBACK:
                        mf = min(mf, Sf)
                        mc = min(mc, Sc)
                dWc = 0
                dtc = 0
                tf = tau
                tc = tau
                tau += d.expo(1 / lam)                              # :474
            dt = tn - tf                                            # :477
            dtc += dt
            dWf = d.normal_w() * np.sqrt(dt)                        # :479
            dWc += dWf
            if functional == ASIAN:                                 # :554-556
                avg_f += 0.5 * dt * Sf
                Sf += merton_increment(p, Sf, dt, dWf)
                avg_f += 0.5 * dt * Sf
            else:
                Sf += merton_increment(p, Sf, dt, dWf)              # :481 / :635
            tf = tn
            if functional == LOOKBACK:
                mf = min(mf, Sf)                                    # :637
            if j % M == 0:
                if functional == ASIAN:                             # :560-562
                    avg_c += 0.5 * dtc * Sc
                    Sc += merton_increment(p, Sc, dtc, dWc)
                    avg_c += 0.5 * dtc * Sc
                else:
                    Sc += merton_increment(p, Sc, dtc, dWc)         # :484 / :639
                    if functional == LOOKBACK:
                        mc = min(mc, Sc)                            # :640
                tc = tn
                dtc = 0
                dWc = 0
        if functional == ASIAN:
            out_f[num] = avg_f
            out_c[num] = avg_c
        else:
            out_f[num] = Sf
            out_c[num] = Sc
            min_f[num] = mf
            min_c[num] = mc
        d.next_path()
    if functional == ASIAN:
        return out_f / T, out_c / T                                 # :571
    if functional == LOOKBACK:
        return out_f, min_f, out_c, min_c
    return out_f, out_c


# ----------------------------------------------------------------------------- payoffs
def functional_of(payoff):
    return {EURO: EURO, DIGITAL: EURO, ASIAN: ASIAN, LOOKBACK: LOOKBACK}[payoff]


def apply_payoff(p, payoff, l, M, path_out):
    """EA_payoff :62-86, Lookback_payoff :89-112, Digital_payoff :114-136.
    At l==0 the second return is the raw coarse path output, as in the reference."""
    r, T, K = p.r, p.T, p.K
    if payoff in (EURO, ASIAN):
        Xf, Xc = path_out
        Pf = np.maximum(0, Xf - K)
        Pf = np.exp(-r * T) * Pf
        if l == 0:
            return Pf, Xc
        Pc = np.maximum(0, Xc - K)
        Pc = np.exp(-r * T) * Pc
        return Pf, Pc
    if payoff == LOOKBACK:
        dt = T / M ** l
        Xf, Mf, Xc, Mc = path_out
        Pf = Xf - Mf * (1 - p.beta * p.sig * np.sqrt(dt))
        Pf = np.exp(-r * T) * Pf
        if l == 0:
            return Pf, Xc
        Pc = Xc - Mc * (1 - p.beta * p.sig * np.sqrt(M * dt))
        Pc = np.exp(-r * T) * Pc
        return Pf, Pc
    if payoff == DIGITAL:
        Xf, Xc = path_out
        Pf = np.exp(-r * T) * K * (Xf > K).astype(np.int_)
        if l == 0:
            return Pf, Xc
        Pc = np.exp(-r * T) * K * (Xc > K).astype(np.int_)
        return Pf, Pc
    raise ValueError(payoff)


def level_payoffs(p, model, payoff, l, M, n=None, Z=None, draws=None):
    """`Option.payoff(N_loop,l,M)` for one of the 8 products (+ American put) -> (Pf, Pc)."""
    if payoff == AMERICAN:
        return amer_level_payoffs(p, l, M, n=n, Z=Z)
    fn = functional_of(payoff)
    if model == GBM:
        out = gbm_paths(p, fn, l, M, Z=Z, n=n)
    else:
        out = merton_paths(p, fn, l, M, n, draws=draws)
    return apply_payoff(p, payoff, l, M, out)


# ----------------------------------------------------------------------------- looper
def seven_sums(Pf, Pc, l):
    """One chunk of Option.looper :946-955."""
    sumPf = np.sum(Pf)
    sumPf2 = np.sum(Pf ** 2)
    if l == 0:                       # Pc is ignored here (an array of X0, or the scalar 0 of Amer_payoff)
        return np.array([sumPf, sumPf2, sumPf, sumPf2, 0, 0, 0])
    dP = Pf - Pc
    return np.array([np.sum(dP), np.sum(dP ** 2), sumPf, sumPf2,
                     np.sum(Pc), np.sum(Pc ** 2), np.sum(Pc * Pf)])


def looper(p, model, payoff, Nl, l, M, Npl=10 ** 4, anti=False):
    """Option.looper :919-957 on the numpy global RNG."""
    num_rem = Nl
    suml = np.zeros(7)
    while num_rem > 0:
        N_loop = min(Npl, num_rem)
        num_rem -= N_loop
        if anti:
            Pf, Pc = anti_level_payoffs(p, payoff, l, M, n=N_loop)          # :942
        else:
            Pf, Pc = level_payoffs(p, model, payoff, l, M, n=N_loop)
        suml += seven_sums(Pf, Pc, l)
    return suml
