/* C restatement of MLMC_RSCAM's per-level coupled sampler.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline.  The product
 * (libmlmcb.so, CUDA) never links or calls it.
 *
 * Parity status: PINNED through oracle/np_oracle.py, which is bit-identical to the unmodified
 * reference for a given np.random.seed (tests/test_oracle_vs_reference.py) and to the golden
 * fixtures in tests/golden/ generated from the reference.  This file is checked against
 * np_oracle on replayed draws (bit-exact for GBM: IEEE double, same operation order, built
 * with -ffp-contract=off; <= 4 ulp for Merton because libm exp() and numpy's exp may differ
 * in the last bit).
 *
 * Citations ":NNN" are line numbers of /root/reference/MLMC_RSCAM.py.  Expression order is
 * the reference's left-to-right Python evaluation order, kept on purpose.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    double X0, K, T, r, sig, drift;          /* Option :796-813, GBM_Option :1330-1347      */
    double lam, jumpmean, jumpstd;           /* Merton_Option :1273-1296                    */
    double beta;                             /* Lookback offset constant :1443 / :1691      */
    double disc;                             /* np.exp(-r*T) as evaluated by the caller :80 */
    double J_bar;                            /* np.exp(jumpmean+0.5*jumpstd**2)-1 :1291     */
    double milstein;                         /* != 0: the notebook's milstein_GBM / milstein_Merton increments */
    double boundary_c;                       /* American put: exercise boundary K*(1-c*sqrt(T-t)) (notebook exb) */
} orc_params;

enum { ORC_GBM = 0, ORC_MERTON = 1 };
enum { ORC_EURO = 0, ORC_ASIAN = 1, ORC_LOOKBACK = 2, ORC_DIGITAL = 3, ORC_AMERICAN = 4 };

static int64_t ipow(int M, int l) { int64_t n = 1; for (int i = 0; i < l; ++i) n *= M; return n; }

/* Scheme switch for the increments below; set per call from orc_params.milstein (the replay and
 * level functions are not re-entrant across schemes, which the tests do not need). */
static __thread int g_milstein = 0;
static __thread double g_sig2 = 0.0;
static void set_scheme(const orc_params *p) { g_milstein = p->milstein != 0.0; g_sig2 = pow(p->sig, 2.0); }

/* GBM_Option.sde :1363  mu*X*dt + sig*X*dW ; notebook cell 1 milstein_GBM adds 0.5*X*(sig**2)*(dW**2-dt) */
static inline double gbm_inc(double mu, double sig, double X, double dt, double dW) {
    if (g_milstein) return mu * X * dt + sig * X * dW + 0.5 * X * g_sig2 * (dW * dW - dt);
    return mu * X * dt + sig * X * dW;
}
/* Merton_Option.sde :1314-1315 ; notebook cell 1 milstein_Merton */
static inline double mjd_inc(double c, double sig, double X, double dt, double dW, double dJ) {
    double dx = g_milstein ? c * X * dt + sig * X * dW + 0.5 * X * g_sig2 * (dW * dW - dt)
                           : c * X * dt + sig * X * dW;
    return dx + (dx + X) * dJ;
}

/* EA_payoff :79-85, Lookback_payoff :105-111, Digital_payoff :131-135 */
static inline double payoff_of(const orc_params *p, int payoff, double X, double Mn, double look_c) {
    switch (payoff) {
    case ORC_EURO:
    case ORC_ASIAN: { double v = X - p->K; v = v > 0 ? v : 0; return p->disc * v; }
    case ORC_LOOKBACK: return p->disc * (X - Mn * look_c);
    default: return p->disc * p->K * (X > p->K ? 1.0 : 0.0);
    }
}

/* A normal source: either a replayed array with a stride, or an internal generator. */
typedef struct { uint64_t s[4]; int have; double spare; } orc_rng;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(orc_rng *g, uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0xD1342543DE82EF95ull + stream;
    for (int i = 0; i < 4; ++i) g->s[i] = splitmix(&x);
    g->have = 0;
}
static inline uint64_t rng_next(orc_rng *g) {   /* xoshiro256++ */
    uint64_t *s = g->s, r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static inline double rng_u01(orc_rng *g) { return ((rng_next(g) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double rng_normal(orc_rng *g) {   /* Box-Muller, both branches used */
    if (g->have) { g->have = 0; return g->spare; }
    double u1 = rng_u01(g), u2 = rng_u01(g);
    double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925 * u2;
    g->spare = r * sin(a); g->have = 1;
    return r * cos(a);
}

/* ---- one coupled GBM path: diffusion_path :320-349, _avg :351-386, _min :388-426 -------- */
static void gbm_one(const orc_params *p, int payoff, int l, int M, int64_t nsteps,
                    const double *Z, int64_t zstride, orc_rng *g, double *Pf, double *Pc) {
    const double T = p->T, X0 = p->X0, mu = p->r + p->drift, sig = p->sig;
    const double dt = T / (double)nsteps, sqrt_dt = sqrt(dt);
    double Xf = X0, Xc = X0, dWc = 0.0;
    double Af = 0.5 * dt * Xf, Ac = 0.5 * M * dt * Xc;             /* :369-370 */
    double Mf = X0, Mc = X0;                                       /* :411-412 */
    for (int64_t j = 1; j <= nsteps; ++j) {
        double z = Z ? Z[(j - 1) * zstride] : rng_normal(g);
        double dWf = z * sqrt_dt;                                  /* :343 */
        dWc = dWc + dWf;                                           /* :344 */
        Xf += gbm_inc(mu, sig, Xf, dt, dWf);                       /* :345 */
        if (payoff == ORC_ASIAN) Af += Xf * dt;                    /* :377 */
        if (payoff == ORC_LOOKBACK) Mf = Xf < Mf ? Xf : Mf;        /* :420 */
        if (j % M == 0) {
            Xc += gbm_inc(mu, sig, Xc, M * dt, dWc);               /* :347 */
            if (payoff == ORC_ASIAN) Ac += Xc * M * dt;            /* :380 */
            if (payoff == ORC_LOOKBACK) Mc = Xc < Mc ? Xc : Mc;    /* :423 */
            dWc = 0.0;                                             /* :348 */
        }
    }
    double vf = Xf, vc = Xc;
    if (payoff == ORC_ASIAN) {
        Af -= 0.5 * Xf * dt;                                       /* :383 */
        Ac -= 0.5 * Xc * M * dt;                                   /* :384 */
        vf = Af / T; vc = Ac / T;                                  /* :386 */
    }
    double cf = 1 - p->beta * sig * sqrt(dt), cc = 1 - p->beta * sig * sqrt(M * dt);   /* :105 :110 */
    *Pf = payoff_of(p, payoff, vf, Mf, cf);
    *Pc = (l == 0) ? vc : payoff_of(p, payoff, vc, Mc, cc);        /* l==0: raw coarse output :82 */
}

/* ---- one coupled American-put path: Amer_GBM.path :1511-1615 + Amer_payoff :138-158 (M == 2) ---- *
 * `halfdWc` aliases `dWc` in the reference (:1577 binds the array that :1549 updates in place), so  *
 * the coarse mid-point uses dWc - 0.5*dWc.                                                          */
static inline double amer_eb(const orc_params *p, double t) {          /* notebook cell 1: exb */
    double v = p->K * (1 - p->boundary_c * sqrt(p->T - t));
    return v > 0 ? v : 0;
}
static inline double pos(double v) { return v > 0 ? v : 0; }
static void amer_one(const orc_params *p, int l, int M, int64_t nsteps, const double *Z, int64_t zstride,
                     orc_rng *g, double *Pf, double *Pc) {
    const double T = p->T, X0 = p->X0, mu = p->r + p->drift, sig = p->sig, K = p->K, r = p->r;
    const double dt = T / (double)nsteps, sqrt_dt = sqrt(dt), sig2 = pow(sig, 2.0);
    double f0 = X0, f1 = 1.0, f2 = 0.0, c0 = X0, c1 = 1.0, c2 = 0.0, dWc = 0.0;
    for (int64_t j = 1; j <= nsteps; ++j) {
        double t_ = (double)(j - 1) * dt;
        double z = Z ? Z[(j - 1) * zstride] : rng_normal(g);
        double dWf = sqrt_dt * z;                                                  /* :1548 */
        dWc += dWf;
        double Xl = f0, Xr = Xl + gbm_inc(mu, sig, Xl, dt, dWf);                   /* :1554 */
        double lb = amer_eb(p, t_), tfMid = t_ + dt / 2, rb = amer_eb(p, (double)j * dt), mb = amer_eb(p, tfMid);
        double P1 = exp(-2 * pos(Xl - lb) * pos(Xr - rb) / (sig2 * dt * (Xl * Xl)));       /* :1566 */
        f2 += f1 * P1 * pos(K - mb) * exp(-r * tfMid);                             /* :1569 */
        f1 *= (1.0 - P1);
        f0 = Xr;
        if (j % M == 0) {
            t_ = (double)(j - M) * dt;
            Xl = c0; Xr = Xl + gbm_inc(mu, sig, Xl, M * dt, dWc);                  /* :1584 */
            double Xmid = Xl + 0.5 * (Xr - Xl) + sig * Xl * (dWc - 0.5 * dWc);     /* :1585 with the alias */
            lb = amer_eb(p, t_);
            double tcMid = t_ + M * dt / 2;
            rb = amer_eb(p, (double)j * dt);
            double mb2 = amer_eb(p, tcMid), mb1 = 0.5 * (rb + lb);
            double P11 = exp(-2 * pos(Xl - lb) * pos(Xmid - mb1) / (dt * sig2 * (Xl * Xl)));     /* :1600 */
            double P12 = exp(-2 * pos(Xmid - mb1) * pos(Xr - rb) / (dt * sig2 * (Xl * Xl)));     /* :1601 */
            double Pc1 = (1.0 - (1.0 - P11) * (1.0 - P12));
            c2 += c1 * Pc1 * pos(K - mb2) * exp(-r * tcMid);                       /* :1605 */
            c1 *= (1 - Pc1);
            c0 = Xr;
            dWc = 0.0;
        }
    }
    *Pf = f2 + f1 * pos(K - f0) * p->disc;                                         /* :153 */
    *Pc = (l == 0) ? 0.0 : c2 + c1 * pos(K - c0) * p->disc;                        /* :155 / :157 */
}

/* ---- one antithetic GBM triple: diffusion_anti_path :162-217, _avg :219-290, anti_EA_payoff :292-317 -- *
 * Returns Pf' = 0.5*(Pf+Pa) and Pc; at l==0 Pf and the raw coarse value (:311).                      */
static void gbm_anti_one(const orc_params *p, int payoff, int l, int M, int64_t nsteps,
                         const double *Z, int64_t zstride, orc_rng *g, double *Pf, double *Pc) {
    const double T = p->T, X0 = p->X0, mu = p->r + p->drift, sig = p->sig;
    const double dt = T / (double)nsteps, sqrt_dt = sqrt(dt);
    double Xf = X0, Xc = X0, Xa = X0, dWc = 0.0;
    double Af = 0.5 * dt * Xf, Aa = 0.5 * dt * Xa, Ac = 0.5 * M * dt * Xc;      /* :247-249 */
    int64_t k = 0;
    for (int64_t j = 2; j <= nsteps; j += 2) {
        double zo = Z ? Z[k * zstride] : rng_normal(g), ze = Z ? Z[(k + 1) * zstride] : rng_normal(g);
        k += 2;
        double dWo = zo * sqrt_dt, dWe = ze * sqrt_dt;                            /* :195-196 */
        Xf += gbm_inc(mu, sig, Xf, dt, dWo);                                      /* :200 */
        Xa += gbm_inc(mu, sig, Xa, dt, dWe);                                      /* :201 */
        if (payoff == ORC_ASIAN) { Af += dt * Xf; Aa += dt * Xa; }                /* :263-264 */
        Xf += gbm_inc(mu, sig, Xf, dt, dWe);                                      /* :205 */
        Xa += gbm_inc(mu, sig, Xa, dt, dWo);                                      /* :206 */
        if (payoff == ORC_ASIAN) { Af += dt * Xf; Aa += dt * Xa; }                /* :271-272 */
        dWc += dWo + dWe;                                                         /* :208 */
        if (j % M == 0) {
            Xc += gbm_inc(mu, sig, Xc, M * dt, dWc);                              /* :210 */
            if (payoff == ORC_ASIAN) Ac += dt * M * Xc;                           /* :278 */
            dWc = 0.0;
        }
    }
    double vf, va, vc;
    if (l == 0) {
        double z = Z ? Z[0] : rng_normal(g);
        Xf += gbm_inc(mu, sig, Xf, dt, z * sqrt_dt);                              /* :214 / :283 */
        if (payoff == ORC_ASIAN) { Af += 0.5 * dt * Xf; vf = Af / T; } else vf = Xf;   /* :284 */
        *Pf = payoff_of(p, payoff, vf, 0, 0);
        *Pc = Xc;
        return;
    }
    if (payoff == ORC_ASIAN) {
        Af -= 0.5 * Xf * dt; Aa -= 0.5 * Xa * dt; Ac -= 0.5 * Xc * (M * dt);      /* :287-289 */
        vf = Af / T; va = Aa / T; vc = Ac / T;
    } else { vf = Xf; va = Xa; vc = Xc; }
    *Pf = 0.5 * (payoff_of(p, payoff, vf, 0, 0) + payoff_of(p, payoff, va, 0, 0));    /* :317 */
    *Pc = payoff_of(p, payoff, vc, 0, 0);
}

/* ---- one coupled Merton path: JD_path :428-492, _avg :494-571, _min :573-651 ------------ */
typedef struct { const double *zW, *zQ, *E; int64_t iw, iq, ie; orc_rng *g; double scale; } jd_src;
static inline double src_w(jd_src *s) { return s->zW ? s->zW[s->iw++] : rng_normal(s->g); }
static inline double src_q(jd_src *s) { return s->zQ ? s->zQ[s->iq++] : rng_normal(s->g); }
static inline double src_e(jd_src *s) { return s->E ? s->E[s->ie++] : -log(rng_u01(s->g)) * s->scale; }

static void mjd_one(const orc_params *p, int payoff, int l, int M, int64_t nsteps, jd_src *s,
                    double *Pf, double *Pc) {
    const double T = p->T, X0 = p->X0, sig = p->sig, c = p->r - p->lam * p->J_bar;
    double dWc = 0, tau = 0, tf = 0, tc = 0, dtc = 0, Sf = X0, Sc = X0;
    double avg_f = 0, avg_c = 0, mf = X0, mc = X0;
    tau += src_e(s);                                               /* :454 */
    for (int64_t j = 1; j <= nsteps; ++j) {
        double tn = (double)j * T / (double)nsteps;                /* :456 */
        while (tau < tn) {                                         /* :457 */
            double dt = tau - tf;
            dtc += dt;
            double dWf = src_w(s) * sqrt(dt);                      /* :460 */
            dWc += dWf;
            double dJ = exp(p->jumpmean + p->jumpstd * src_q(s)) - 1;      /* :462 */
            if (payoff == ORC_ASIAN) {                             /* :532-540 */
                avg_f += 0.5 * dt * Sf;  Sf += mjd_inc(c, sig, Sf, dt, dWf, 0.0);
                avg_f += 0.5 * dt * Sf;  Sf += Sf * dJ;
                avg_c += 0.5 * dtc * Sc; Sc += mjd_inc(c, sig, Sc, dtc, dWc, 0.0);
                avg_c += 0.5 * dtc * Sc; Sc += Sc * dJ;
            } else {                                               /* :465-468 */
                Sf += mjd_inc(c, sig, Sf, dt, dWf, dJ);
                Sc += mjd_inc(c, sig, Sc, dtc, dWc, dJ);
                if (payoff == ORC_LOOKBACK) { mf = Sf < mf ? Sf : mf; mc = Sc < mc ? Sc : mc; }  /* :621-622 */
            }
            dWc = 0; dtc = 0; tf = tau; tc = tau;                  /* :470-473 */
            tau += src_e(s);                                       /* :474 */
        }
        double dt = tn - tf;                                       /* :477 */
        dtc += dt;
        double dWf = src_w(s) * sqrt(dt);                          /* :479 */
        dWc += dWf;
        if (payoff == ORC_ASIAN) {                                 /* :554-556 */
            avg_f += 0.5 * dt * Sf; Sf += mjd_inc(c, sig, Sf, dt, dWf, 0.0); avg_f += 0.5 * dt * Sf;
        } else {
            Sf += mjd_inc(c, sig, Sf, dt, dWf, 0.0);               /* :481 */
        }
        tf = tn;
        if (payoff == ORC_LOOKBACK) mf = Sf < mf ? Sf : mf;        /* :637 */
        if (j % M == 0) {
            if (payoff == ORC_ASIAN) {                             /* :560-562 */
                avg_c += 0.5 * dtc * Sc; Sc += mjd_inc(c, sig, Sc, dtc, dWc, 0.0); avg_c += 0.5 * dtc * Sc;
            } else {
                Sc += mjd_inc(c, sig, Sc, dtc, dWc, 0.0);          /* :484 */
                if (payoff == ORC_LOOKBACK) mc = Sc < mc ? Sc : mc;        /* :640 */
            }
            tc = tn; dtc = 0; dWc = 0;
        }
    }
    (void)tc;
    double vf = Sf, vc = Sc;
    if (payoff == ORC_ASIAN) { vf = avg_f / T; vc = avg_c / T; }   /* :571 */
    double dtg = T / (double)nsteps;                               /* Lookback_payoff :102 */
    double cf = 1 - p->beta * sig * sqrt(dtg), cc = 1 - p->beta * sig * sqrt(M * dtg);
    *Pf = payoff_of(p, payoff, vf, mf, cf);
    *Pc = (l == 0) ? vc : payoff_of(p, payoff, vc, mc, cc);
}

/* ---- replay entry points ---------------------------------------------------------------- */
int orc_gbm_replay(const orc_params *p, int payoff, int l, int M, int64_t n,
                   const double *Z /* [M**l][n], step-major as np.random.randn(N_loop) per step */,
                   double *Pf, double *Pc) {
    int64_t nsteps = ipow(M, l);
    set_scheme(p);
    if (payoff == ORC_AMERICAN) {
        if (M != 2) return -1;
        for (int64_t i = 0; i < n; ++i) amer_one(p, l, M, nsteps, Z + i, n, NULL, Pf + i, Pc + i);
        return 0;
    }
    for (int64_t i = 0; i < n; ++i) gbm_one(p, payoff, l, M, nsteps, Z + i, n, NULL, Pf + i, Pc + i);
    return 0;
}

int orc_gbm_anti_replay(const orc_params *p, int payoff, int l, int M, int64_t n, const double *Z,
                        double *Pf, double *Pc) {
    if (M % 2 || (payoff != ORC_EURO && payoff != ORC_ASIAN)) return -1;
    int64_t nsteps = ipow(M, l);
    set_scheme(p);
    for (int64_t i = 0; i < n; ++i) gbm_anti_one(p, payoff, l, M, nsteps, Z + i, n, NULL, Pf + i, Pc + i);
    return 0;
}

int orc_merton_replay(const orc_params *p, int payoff, int l, int M, int64_t n,
                      const double *zW, const int64_t *offW, const double *zQ, const int64_t *offQ,
                      const double *E, const int64_t *offE, double *Pf, double *Pc) {
    int64_t nsteps = ipow(M, l);
    set_scheme(p);
    for (int64_t i = 0; i < n; ++i) {
        jd_src s = { zW, zQ, E, offW[i], offQ[i], offE[i], NULL, 1.0 / p->lam };
        mjd_one(p, payoff, l, M, nsteps, &s, Pf + i, Pc + i);
    }
    return 0;
}

/* ---- Option.looper :919-957 with an internal generator (xoshiro256++ / Box-Muller) ------ *
 * Statistical twin of the reference level function for sample counts the Python loops cannot
 * reach, and the multi-threaded "C port" CPU baseline (pthreads; this image's gcc has no
 * libgomp).  out7 row order :955; l==0 row :949.  Thread t takes a contiguous slice of path
 * ids and the partial rows are added in thread order, so the result is deterministic.      */
typedef struct {
    const orc_params *p; int model, payoff, l, M, anti; int64_t nsteps, lo, hi; uint64_t seed; double a[7];
} orc_job;

static void *orc_worker(void *arg) {
    orc_job *jb = (orc_job *)arg;
    double *a = jb->a;
    set_scheme(jb->p);
    for (int k = 0; k < 7; ++k) a[k] = 0;
    for (int64_t i = jb->lo; i < jb->hi; ++i) {
        orc_rng g; rng_seed(&g, jb->seed, (uint64_t)i + ((uint64_t)jb->l << 48));
        double Pf, Pc;
        if (jb->payoff == ORC_AMERICAN) amer_one(jb->p, jb->l, jb->M, jb->nsteps, NULL, 0, &g, &Pf, &Pc);
        else if (jb->anti) gbm_anti_one(jb->p, jb->payoff, jb->l, jb->M, jb->nsteps, NULL, 0, &g, &Pf, &Pc);
        else if (jb->model == ORC_GBM) gbm_one(jb->p, jb->payoff, jb->l, jb->M, jb->nsteps, NULL, 0, &g, &Pf, &Pc);
        else {
            jd_src s = { NULL, NULL, NULL, 0, 0, 0, &g, 1.0 / jb->p->lam };
            mjd_one(jb->p, jb->payoff, jb->l, jb->M, jb->nsteps, &s, &Pf, &Pc);
        }
        if (jb->l == 0) { a[0] += Pf; a[1] += Pf * Pf; a[2] += Pf; a[3] += Pf * Pf; }
        else {
            double d = Pf - Pc;
            a[0] += d; a[1] += d * d; a[2] += Pf; a[3] += Pf * Pf; a[4] += Pc; a[5] += Pc * Pc; a[6] += Pc * Pf;
        }
    }
    return NULL;
}

static int level_sums_impl(const orc_params *p, int model, int payoff, int l, int M, int64_t n,
                           uint64_t seed, int nthreads, int anti, double *out7) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    orc_job jobs[256]; pthread_t th[256];
    int64_t nsteps = ipow(M, l);
    for (int t = 0; t < nthreads; ++t) {
        orc_job jb = { p, model, payoff, l, M, anti, nsteps, n * t / nthreads, n * (t + 1) / nthreads, seed, {0} };
        jobs[t] = jb;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    orc_worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    for (int k = 0; k < 7; ++k) { double v = 0; for (int t = 0; t < nthreads; ++t) v += jobs[t].a[k]; out7[k] = v; }
    return 0;
}

int orc_level_sums(const orc_params *p, int model, int payoff, int l, int M, int64_t n,
                   uint64_t seed, int nthreads, double *out7) {
    return level_sums_impl(p, model, payoff, l, M, n, seed, nthreads, 0, out7);
}

/* looper(..., anti=True) :941-942 for Euro_GBM / Asian_GBM */
int orc_level_sums_anti(const orc_params *p, int payoff, int l, int M, int64_t n,
                        uint64_t seed, int nthreads, double *out7) {
    if (M % 2 || (payoff != ORC_EURO && payoff != ORC_ASIAN)) return -1;
    return level_sums_impl(p, ORC_GBM, payoff, l, M, n, seed, nthreads, 1, out7);
}

int orc_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
