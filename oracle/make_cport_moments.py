"""Level moments of the pinned C restatement (oracle/c_oracle.c, own generator) at sample counts the
Python reference cannot reach: 10^7 coupled samples per case at the headline level l=6, M=4 (GBM) and at
l=7, M=2 (Merton, 128 fine steps).  TEST INFRASTRUCTURE: the -m gpu statistical-parity tests compare the
CUDA kernels' level means and variances with these at 3 standard errors (north star (b)).

    python oracle/make_cport_moments.py          # ~10 minutes on 8 cores -> tests/golden/cport_moments.json

The C port is bit-exact against the reference's fixtures for GBM and within 2e-13*X0 for Merton
(tests/test_oracle_golden.py), so its moments are the reference algorithm's moments."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, np_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cport_moments.json")
BENCH = dict(X0=125, K=105, r=0.02, sig=0.5)
CASES = [
    # (name, class name, model, payoff, ctor kwargs, scheme, l, M, n)
    ("euro_gbm_bench_l6M4", "Euro_GBM", O.GBM, O.EURO, BENCH, "em", 6, 4, 10_000_000),
    ("euro_gbm_l6M4", "Euro_GBM", O.GBM, O.EURO, {}, "em", 6, 4, 10_000_000),
    ("asian_gbm_l6M4", "Asian_GBM", O.GBM, O.ASIAN, {}, "em", 6, 4, 10_000_000),
    ("lookback_gbm_l6M4", "Lookback_GBM", O.GBM, O.LOOKBACK, {}, "em", 6, 4, 10_000_000),
    ("digital_gbm_l6M4", "Digital_GBM", O.GBM, O.DIGITAL, {}, "em", 6, 4, 10_000_000),
    ("euro_gbm_milstein_l4M4", "Euro_GBM", O.GBM, O.EURO, {}, "milstein", 4, 4, 20_000_000),
    ("euro_merton_l7M2", "Euro_Merton", O.MERTON, O.EURO, {}, "em", 7, 2, 10_000_000),
    ("asian_merton_l7M2", "Asian_Merton", O.MERTON, O.ASIAN, {}, "em", 7, 2, 10_000_000),
    ("lookback_merton_l7M2", "Lookback_Merton", O.MERTON, O.LOOKBACK, {}, "em", 7, 2, 10_000_000),
    ("digital_merton_l7M2", "Digital_Merton", O.MERTON, O.DIGITAL, {}, "em", 7, 2, 10_000_000),
    ("euro_merton_l4M2", "Euro_Merton", O.MERTON, O.EURO, {}, "em", 4, 2, 40_000_000),
    ("digital_merton_l2M2", "Digital_Merton", O.MERTON, O.DIGITAL, {}, "em", 2, 2, 40_000_000),
]


def main():
    c_oracle.build()
    threads = os.cpu_count() or 1
    res = {}
    if os.path.exists(OUT):
        res = json.load(open(OUT))
    for i, (name, cname, model, pay, kw, scheme, l, M, n) in enumerate(CASES):
        if name in res and res[name]["n"] == n:
            continue
        t0 = time.time()
        seed = 90_000 + i
        s = c_oracle.level_sums(O.Params(scheme=scheme, **kw), model, pay, l, M, n, seed=seed, nthreads=threads)
        res[name] = dict(cls=cname, kwargs=kw, scheme=scheme, l=l, M=M, n=n, seed=seed, sums=[float(x) for x in s])
        print(f"{name}: {time.time() - t0:.1f} s  E[dP]={s[0] / n:.6g} E[Pf]={s[2] / n:.6g}", flush=True)
        json.dump(res, open(OUT, "w"), indent=1)
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
