"""Plot-free numeric core of the reference's `Giles_plot` (MLMC_RSCAM.py:1708-1891).

`giles_analysis` computes exactly what the four panels of the reference figure display - per-eps
sample allocations N_l and eps^2*cost for MLMC / plain MC / antithetic MLMC (:1741-1772), the
fixed-level sweep of means and variances of P_l and P_l - P_{l-1} (:1775-1782, antithetic
:1798-1805) and the regressed orders alpha, beta (:1818-1823) - and returns them as arrays instead
of drawing them.  `option` may be one of this package's classes (the sweep goes down as one batched
GPU call) or any object with the reference's `mlmc` / `looper` methods.
"""
import numpy as np


def _level_sums(option, Lmax, Nsamples, M, anti):
    if hasattr(option, "level_sweep"):
        return option.level_sweep(Lmax=Lmax, Nsamples=Nsamples, M=M, anti=anti)[:4]
    sums = np.zeros((4, Lmax + 1))
    for l in range(Lmax + 1):                                   # :1776-1777 / :1799-1800
        sums[:, l] = option.looper(Nsamples, l, M, anti=anti)[:4]
    return sums


def giles_analysis(option, eps, M=2, N0=10 ** 3, anti=False, Lmax=8, Nsamples=10 ** 6):
    eps = list(eps)
    cost_mlmc, cost_mc, cost_anti = [], [], []
    N_std, N_anti = [], []
    for e in eps:                                               # :1741-1772
        sums, N = option.mlmc(e, M, warm_start=False, N0=N0)
        L = len(N) - 1
        means_p = np.abs(sums[2, :] / N)
        V_p = (sums[3, :] / N) - means_p ** 2
        # cost = calls to the rng, i.e. fine steps (+ jumps for jump diffusions)  :1748-1755
        if hasattr(option, 'lam'):
            cost_mlmc += [(np.sum(N) * option.lam + np.sum(N * (M ** np.arange(0, L + 1)))) * e ** 2]
            cost_mc += [2 * np.sum(V_p * (M ** np.arange(L + 1) + option.lam))]
        else:
            cost_mlmc += [np.sum(N * (M ** np.arange(0, L + 1))) * e ** 2]
            cost_mc += [2 * np.sum(V_p * (M ** np.arange(L + 1)))]
        N_std.append(np.array(N, copy=True))
        if anti == True:  # noqa: E712                          # :1760-1769
            sums, N = option.mlmc(e, M, anti=anti, warm_start=False, N0=N0)
            L = len(N) - 1
            if hasattr(option, 'lam'):
                cost_anti += [(np.sum(N) * option.lam + np.sum(N * (M ** np.arange(0, L + 1)))) * e ** 2]
            else:
                cost_anti += [np.sum(N * (M ** np.arange(0, L + 1))) * e ** 2]
            N_anti.append(np.array(N, copy=True))

    sums = _level_sums(option, Lmax, Nsamples, M, False)        # :1775-1777
    means_p = np.abs(sums[2, :] / Nsamples)
    V_p = (sums[3, :] / Nsamples) - means_p ** 2
    means_dp = np.abs(sums[0, :] / Nsamples)
    V_dp = (sums[1, :] / Nsamples) - means_dp ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        out = {
            "eps": eps, "M": M, "Lmax": Lmax, "Nsamples": Nsamples,
            "N": N_std, "N_anti": N_anti,
            "cost_mlmc": np.array(cost_mlmc), "cost_mc": np.array(cost_mc), "cost_anti": np.array(cost_anti),
            "log_var_P": np.log(V_p) / np.log(M), "log_var_dP": np.log(V_dp[1:]) / np.log(M),          # :1785-1787
            "log_mean_P": np.log(means_p) / np.log(M), "log_mean_dP": np.log(means_dp[1:]) / np.log(M),  # :1791-1794
        }
    if anti == True:  # noqa: E712                              # :1798-1805 (overwrites the four vectors)
        sums = _level_sums(option, Lmax, Nsamples, M, True)
        means_p = np.abs(sums[2, :] / Nsamples)
        V_p = (sums[3, :] / Nsamples) - means_p ** 2
        means_dp = np.abs(sums[0, :] / Nsamples)
        V_dp = (sums[1, :] / Nsamples) - means_dp ** 2
        out["log_var_dP_anti"] = np.log(V_dp[1:]) / np.log(M)
        out["log_mean_dP_anti"] = np.log(means_dp[1:]) / np.log(M)

    X = np.ones((Lmax, 2))                                      # :1818-1823 (anti values if anti)
    X[:, 0] = np.arange(1, Lmax + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.linalg.lstsq(X, np.log(means_dp[1:]), rcond=None)[0]
        out["alpha"] = -a[0] / np.log(M)
        b = np.linalg.lstsq(X, np.log(V_dp[1:]), rcond=None)[0]
        out["beta"] = -b[0] / np.log(M)
    return out
