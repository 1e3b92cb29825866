#!/usr/bin/env python
"""Large-sample check of the kernel's normal samplers (evidence for DESIGN.md 2.2).

    python tools/sampler_quality.py [--draws 4e9] > gpurun_out/sampler_quality.json

Streams mlmcb_sampler_probe output through torch reductions on the GPU: raw moments 1..8 against
N(0,1) with their standard errors, a 64-bin chi-square on [-4,4] + tails, tail masses at 3/4/5
sigma, lag-1 autocorrelation along a path and across neighbouring paths.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multilevel_mc_sdes_b200 import _lib  # noqa: E402

NORMAL_MOMENTS = {1: 0, 2: 1, 3: 0, 4: 3, 5: 0, 6: 15, 7: 0, 8: 105}


def double_fact_var(k):
    # Var(z^k) = E z^{2k} - (E z^k)^2,  E z^{2k} = (2k-1)!!
    m2k = math.prod(range(1, 2 * k, 2))
    return m2k - NORMAL_MOMENTS[k] ** 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=float, default=4e9)
    ap.add_argument("--seed", type=int, default=2024)
    a = ap.parse_args()
    lib = _lib.load()
    n_paths, n_steps = 1 << 20, 128
    chunk = n_paths * n_steps
    n_chunks = max(1, int(a.draws // chunk))
    buf = torch.empty(chunk, dtype=torch.float64, device="cuda")
    edges = torch.linspace(-4, 4, 65, dtype=torch.float64, device="cuda")
    rep = {}
    for name, samp in (("f32bm", _lib.SAMPLER_F32BM), ("f64bm", _lib.SAMPLER_F64BM)):
        mom = torch.zeros(8, dtype=torch.float64, device="cuda")
        hist = torch.zeros(66, dtype=torch.float64, device="cuda")
        tails = torch.zeros(3, dtype=torch.float64, device="cuda")
        lag_t = torch.zeros(1, dtype=torch.float64, device="cuda")
        lag_p = torch.zeros(1, dtype=torch.float64, device="cuda")
        zmax = 0.0
        chunks = n_chunks if samp == _lib.SAMPLER_F32BM else max(1, n_chunks // 4)
        for c in range(chunks):
            _lib.check(lib.mlmcb_sampler_probe(samp, a.seed, 6, c * n_paths, n_paths, n_steps, 0, 0, buf.data_ptr()))
            z = buf
            p = torch.ones_like(z)
            for k in range(8):
                p = p * z
                mom[k] += p.sum()
            idx = torch.bucketize(z, edges)          # 0 = below -4, 65 = above 4
            hist += torch.bincount(idx, minlength=66).double()
            az = z.abs()
            for i, t in enumerate((3.0, 4.0, 5.0)):
                tails[i] += (az > t).sum()
            zz = z.view(n_steps, n_paths)
            lag_t += (zz[:-1] * zz[1:]).sum()
            lag_p += (zz[:, :-1] * zz[:, 1:]).sum()
            zmax = max(zmax, float(az.max()))
        N = chunks * chunk
        torch.cuda.synchronize()
        out = {"draws": N, "max_abs": zmax, "moments": {}}
        worst = 0.0
        for k in range(1, 9):
            est = float(mom[k - 1]) / N
            se = math.sqrt(double_fact_var(k) / N)
            zscore = (est - NORMAL_MOMENTS[k]) / se
            worst = max(worst, abs(zscore))
            out["moments"][str(k)] = {"estimate": est, "expected": NORMAL_MOMENTS[k], "z": zscore}
        # chi-square over 64 inner bins + 2 tails
        cdf = 0.5 * (1 + torch.erf(edges / math.sqrt(2)))
        pexp = torch.cat([cdf[:1], cdf[1:] - cdf[:-1], 1 - cdf[-1:]])
        chi2 = float((((hist - N * pexp) ** 2) / (N * pexp)).sum())
        out["chi2_66bins"] = {"value": chi2, "dof": 65, "z": (chi2 - 65) / math.sqrt(130)}
        for i, (t, pt) in enumerate(((3, 2.6997960632601866e-3), (4, 6.334248366623996e-5), (5, 5.733031437583869e-7))):
            est = float(tails[i]) / N
            out[f"tail_{t}sigma"] = {"estimate": est, "expected": pt, "z": (est - pt) / math.sqrt(pt / N)}
        out["lag1_along_path_z"] = float(lag_t) / math.sqrt(chunks * (n_steps - 1) * n_paths)
        out["lag1_across_paths_z"] = float(lag_p) / math.sqrt(chunks * n_steps * (n_paths - 1))
        out["worst_moment_z"] = worst
        rep[name] = out
    rep["jump_records"] = jump_records(lib, a)
    print(json.dumps(rep, indent=1))


def jump_records(lib, a):
    """Jump-stream records (mlmcb_jump_probe): E = lam*gap against Exp(1) - raw moments 1..4, a 64-bin
    chi-square on exp(-E) (uniform under the null) + the mass beyond E = 12, extremes - and the two
    normals of the record: moments 1..4, correlations with each other, with E and with the next gap."""
    lam, n_paths, K = 2.5, 1 << 20, 64
    chunk = n_paths * K
    n_chunks = max(1, int(a.draws / 8 // chunk))
    buf = torch.empty(3 * chunk, dtype=torch.float64, device="cuda")
    out = {}
    for name, samp in (("f32bm", _lib.SAMPLER_F32BM), ("f64bm", _lib.SAMPLER_F64BM)):
        chunks = n_chunks if samp == _lib.SAMPLER_F32BM else max(1, n_chunks // 4)
        me = torch.zeros(4, dtype=torch.float64, device="cuda")
        mz = torch.zeros(2, 4, dtype=torch.float64, device="cuda")
        hist = torch.zeros(64, dtype=torch.float64, device="cuda")
        cross = torch.zeros(4, dtype=torch.float64, device="cuda")   # zq*zw, (E-1)*zq, (E-1)*zw, (E_k-1)(E_k+1-1)
        far = torch.zeros(1, dtype=torch.float64, device="cuda")
        emin, emax = float("inf"), 0.0
        for c in range(chunks):
            _lib.check(lib.mlmcb_jump_probe(samp, a.seed + 1, 3, lam, c * n_paths, n_paths, K, 0, 0, buf.data_ptr()))
            E, zq, zw = (buf.view(3, K, n_paths)[i] for i in range(3))
            E = E * lam
            p = torch.ones_like(E)
            for k in range(4):
                p = p * E
                me[k] += p.sum()
            for j, z in enumerate((zq, zw)):
                p = torch.ones_like(z)
                for k in range(4):
                    p = p * z
                    mz[j, k] += p.sum()
            u = torch.exp(-E)
            hist += torch.bincount((u * 64).long().clamp_(0, 63).view(-1), minlength=64).double()
            far += (E > 12.0).sum()
            cross[0] += (zq * zw).sum()
            cross[1] += ((E - 1) * zq).sum()
            cross[2] += ((E - 1) * zw).sum()
            cross[3] += ((E[:-1] - 1) * (E[1:] - 1)).sum()
            emin, emax = min(emin, float(E.min())), max(emax, float(E.max()))
        torch.cuda.synchronize()
        N = chunks * chunk
        r = {"records": N, "lam_gap_min": emin, "lam_gap_max": emax, "gap_moments": {}, "normal_moments": {}}
        for k in range(1, 5):
            exp_k, var_k = math.factorial(k), math.factorial(2 * k) - math.factorial(k) ** 2
            est = float(me[k - 1]) / N
            r["gap_moments"][str(k)] = {"estimate": est, "expected": exp_k, "z": (est - exp_k) / math.sqrt(var_k / N)}
            for j, nm in enumerate(("zq", "zw")):
                est = float(mz[j, k - 1]) / N
                r["normal_moments"][f"{nm}_{k}"] = {"estimate": est, "z": (est - NORMAL_MOMENTS[k]) / math.sqrt(double_fact_var(k) / N)}
        chi2 = float((((hist - N / 64) ** 2) / (N / 64)).sum())
        r["chi2_64bins_exp_minus_E"] = {"value": chi2, "dof": 63, "z": (chi2 - 63) / math.sqrt(126)}
        pt = math.exp(-12.0)
        r["mass_beyond_12"] = {"estimate": float(far) / N, "expected": pt, "z": (float(far) / N - pt) / math.sqrt(pt / N)}
        r["corr_z"] = {"zq_zw": float(cross[0]) / math.sqrt(N), "E_zq": float(cross[1]) / math.sqrt(N),
                       "E_zw": float(cross[2]) / math.sqrt(N), "E_Enext": float(cross[3]) / math.sqrt(chunks * (K - 1) * n_paths)}
        out[name] = r
    return out


if __name__ == "__main__":
    main()
