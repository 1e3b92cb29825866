/* mlmcb.h - C ABI of libmlmcb.so, the B200 (sm_100a) per-level coupled sampler.
 *
 * Drop-in boundary for ONE path of mckeownish/Multilevel_MC_SDEs (MLMC_RSCAM.py): the level
 * function `Option.looper(Nl,l,M,anti,Npl)` (:919-957) with everything below it
 * (`payoff` :62-136, `(diffusion|JD)_path(_avg|_min)` :320-651, `sde` :1298-1315 / :1349-1363,
 * numpy RNG).  The reference has no FFI of its own (pure Python); these entry points are what
 * a ctypes binding on the reference side would call - see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; `d_` = device pointer (valid on `device`),
 * `h_` = host pointer; every function returns 0 on success or a negative MLMCB_E* code and
 * never throws; mlmcb_last_error() gives the text for the calling thread.  All launches are
 * asynchronous on `cuda_stream` (a cudaStream_t / CUstream, NULL = default stream) unless the
 * name ends in `_host`, which synchronises that stream before returning.
 * There is NO CPU fallback: without a CUDA device every compute call fails with MLMCB_ECUDA.
 */
#ifndef MLMCB_H
#define MLMCB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parameter struct = the attributes the reference classes carry
 * (Option.__init__ :796-813, GBM_Option.__init__ :1330-1347, Merton_Option.__init__ :1273-1296,
 *  Lookback beta :1443/:1691).  `disc` and `J_bar` are the two derived constants the reference
 * evaluates with numpy (np.exp(-r*T) :80, np.exp(jumpmean+0.5*jumpstd**2)-1 :1291); a caller
 * that wants replay results bit-identical to numpy passes numpy's values, anyone else calls
 * mlmcb_params_fill_derived(). */
typedef struct mlmcb_params {
    double X0, K, T, r, sig, drift;
    double lam, jumpmean, jumpstd;
    double beta;
    double disc;
    double J_bar;
    double boundary_c;   /* MLMCB_AMERICAN: exercise boundary b(t) = max(K*(1 - c*sqrt(T-t)), 0) */
    /* MLMCB_AMERICAN, optional: a tabulated exercise boundary for callers whose boundary is code the
     * kernel cannot run (the reference takes any Python callable, Amer_GBM.__init__ :1495-1497).  For a
     * level l < amer_levels with amer_boundary[l] != NULL the n = 2^l-step table is built from these HOST
     * values instead of boundary_c:  b(j*dt), j = 0..n  (:1561/:1568)  |  b((j-1)*dt + dt/2), j = 1..n
     * (:1565/:1571)  |  b((j-2)*dt + M*dt/2), j = 2,4,..,n  (:1593/:1599)  - 2n+1+n/2 doubles, in this
     * order, evaluated by the caller at exactly those times.  NULL / 0: the boundary_c family. */
    const double* const* amer_boundary;
    int amer_levels;
    int reserved_;
} mlmcb_params;

enum { MLMCB_GBM = 0, MLMCB_MERTON = 1 };                                   /* model   */
enum { MLMCB_EURO = 0, MLMCB_ASIAN = 1, MLMCB_LOOKBACK = 2, MLMCB_DIGITAL = 3,       /* payoff */
       MLMCB_AMERICAN = 4 /* American put, Amer_GBM.path :1511-1615 + Amer_payoff :138-158: GBM only, M == 2
                             (:1524), the demo notebook's boundary family (ExampleUsages.ipynb cell 1) */ };

/* flags for the Philox-mode level function */
enum {
    MLMCB_SAMPLER_F64BM = 0,   /* default: 52-bit uniforms, fp64 Box-Muller (table-driven log, Newton sqrt,
                                  table + polynomial sin/cos, <= 1 ulp) - the precision of the reference's
                                  numpy normals                                                        */
    MLMCB_SAMPLER_F32BM = 1,   /* opt-in: 32-bit uniforms, fp32 Box-Muller (MUFU log/sqrt), promoted to fp64;
                                  about twice the throughput, fp32-grade draws (Merton: fp32 Exp() gaps)  */
    MLMCB_SAMPLER_MASK  = 0xF,
    MLMCB_FP32_PATHS    = 0x10,/* opt-in: path arithmetic in fp32 (sums stay fp64), Euler-Maruyama only       */
    MLMCB_MILSTEIN      = 0x40,/* Milstein instead of Euler-Maruyama increments: the reference's demo
                                  notebook swaps `sde` for milstein_GBM / milstein_Merton
                                  (ExampleUsages.ipynb cell 1: + 0.5*X*sig**2*(dW**2-dt)).  Every sampler, with or
                                  without MLMCB_ANTITHETIC; fp64 paths only (with MLMCB_FP32_PATHS -> MLMCB_ENOTIMPL).
                                  May also be ORed into `arith` of the replay entry points.          */
    MLMCB_ANTITHETIC    = 0x20 /* antithetic estimator (looper(anti=True) :941-942): GBM Euro/Asian
                                  only (anti_EA_payoff :292-317, diffusion_anti_path(_avg) :162-290),
                                  even M; other products -> MLMCB_ENOTIMPL, odd M -> MLMCB_EINVAL.
                                  May also be ORed into `arith` of mlmcb_replay_gbm.                */
};
/* arithmetic policy for replay */
enum {
    MLMCB_ARITH_FAST  = 0,     /* production arithmetic (FMA form), the one Philox mode uses       */
    MLMCB_ARITH_EXACT = 1      /* reference expression order, no contraction: bit-exact vs numpy   */
};

enum { MLMCB_OK = 0, MLMCB_EINVAL = -1, MLMCB_ECUDA = -2, MLMCB_ENOTIMPL = -3 };

/* ---- the level function (replaces Option.looper :919-957, anti=False) --------------------
 * Simulates coupled paths with global ids [path_offset, path_offset+n_paths) at level l
 * (M**l fine steps, M**(l-1) coarse) and writes the seven sums in the reference's row order
 * (:955)  [sum dP, sum dP^2, sum Pf, sum Pf^2, sum Pc, sum Pc^2, sum Pf*Pc];  at l==0
 * [sum Pf, sum Pf^2, sum Pf, sum Pf^2, 0, 0, 0] (:949).  Randomness is Philox4x32-10 addressed
 * by (seed, l, global path id, draw index): disjoint id ranges give independent samples, the
 * same id range gives the same samples on any GPU count.  d_Pf / d_Pc (each n_paths doubles)
 * are optional per-path outputs = what Option.payoff (:62-136) returns; pass NULL to skip.
 * The result is deterministic for a given (arguments, GPU model).                           */
int mlmcb_level_sums(const mlmcb_params* p, int model, int payoff, int l, int M,
                     uint64_t n_paths, uint64_t seed, uint64_t path_offset, int flags,
                     int device, void* cuda_stream,
                     double* d_out7, double* d_Pf, double* d_Pc);

/* The path functions themselves (Option.path / anti_path :320-651, :162-290, :1511-1615): same
 * simulation, and additionally the per-path quantities the reference path function returns, as rows
 * of d_raw[rows][n_paths]:  plain products 4 rows (fine value, fine minimum, coarse value, coarse
 * minimum; value = terminal price or average; minima only for lookbacks), MLMCB_ANTITHETIC 3 rows
 * (fine, antithetic, coarse), MLMCB_AMERICAN 6 rows (Xf[:,0..2], Xc[:,0..2]).  mlmcb_path_rows tells. */
int mlmcb_path_rows(int payoff, int flags);
int mlmcb_level_paths(const mlmcb_params* p, int model, int payoff, int l, int M,
                      uint64_t n_paths, uint64_t seed, uint64_t path_offset, int flags,
                      int device, void* cuda_stream, double* d_out7, double* d_raw);

/* Same, result delivered to HOST memory (h_out7[7]); launch + 56-byte D2H + stream sync.
 * This is the call `Option.looper` makes. */
int mlmcb_level_sums_host(const mlmcb_params* p, int model, int payoff, int l, int M,
                          uint64_t n_paths, uint64_t seed, uint64_t path_offset, int flags,
                          int device, void* cuda_stream, double* h_out7);

/* One sweep of Option.mlmc's inner `for l in range(L+1)` (:994-997): n_levels independent
 * level calls in one launch (see mlmcb_level_sweep), one D2H of h_out[n_levels][7].
 * levels[i] is the level index, n_paths[i] may be 0 (row left zero).                        */
int mlmcb_level_sweep_host(const mlmcb_params* p, int model, int payoff, int M, int n_levels,
                           const int* levels, const uint64_t* n_paths, const uint64_t* path_offsets,
                           uint64_t seed, int flags, int device, void* cuda_stream, double* h_out);

/* The same sweep without the host hop: rows of EIGHT doubles (seven sums + one pad) into d_out[n_levels][8]
 * on the device, asynchronous on `cuda_stream`.  A multi-GPU caller all-reduces d_out in place (NCCL) and
 * copies it back once.  Plain estimator on fp64 paths (either scheme, either sampler): ONE
 * kernel launch for all levels - the grid is cut into per-level block ranges, deepest level first - on a
 * persistent workspace (no allocation, no memset); other requests (antithetic, fp32 paths, American put, more
 * than 20 non-empty levels) run as concurrent per-level launches.  Consecutive sweeps on one device are
 * ordered among themselves, whatever streams they are issued on.                                         */
int mlmcb_level_sweep(const mlmcb_params* p, int model, int payoff, int M, int n_levels,
                      const int* levels, const uint64_t* n_paths, const uint64_t* path_offsets,
                      uint64_t seed, int flags, int device, void* cuda_stream, double* d_out);

/* ---- replay mode (parity check (a) of the north star) ------------------------------------
 * Feeds the reference's own draws.  GBM: d_Z[step*n_paths + path] is the np.random.randn(N_loop)
 * vector of step `step+1` (:343).  Merton: three ragged per-path streams in CSR layout, in the
 * order the reference consumes them (:454-479): d_zW dW normals, d_zQ jump-size normals,
 * d_E inter-arrival times (already scaled by 1/lam); d_off* have n_paths+1 entries.
 * Outputs per path (Pf,Pc) as Option.payoff returns them (at l==0 Pc = raw coarse path value,
 * :82) and, if d_out7 != NULL, the seven sums.                                               */
int mlmcb_replay_gbm(const mlmcb_params* p, int payoff, int l, int M, uint64_t n_paths,
                     const double* d_Z, int arith, int device, void* cuda_stream,
                     double* d_Pf, double* d_Pc, double* d_out7);
int mlmcb_replay_merton(const mlmcb_params* p, int payoff, int l, int M, uint64_t n_paths,
                        const double* d_zW, const int64_t* d_offW,
                        const double* d_zQ, const int64_t* d_offQ,
                        const double* d_E, const int64_t* d_offE,
                        int arith, int device, void* cuda_stream,
                        double* d_Pf, double* d_Pc, double* d_out7);

/* ---- test / measurement hooks ------------------------------------------------------------ */
/* The standard normals the Philox-mode kernels use for grid steps: d_out[step*n_paths+path],
 * step < n_steps, for global path ids path_offset.. at level l.  For distribution tests.    */
int mlmcb_sampler_probe(int sampler, uint64_t seed, int l, uint64_t path_offset, uint64_t n_paths,
                        uint64_t n_steps, int device, void* cuda_stream, double* d_out);
/* Same for the coarsest GBM levels (n_steps = 1 or 2 fine steps per path), where neighbouring path ids
 * share one Philox block: d_out[step*n_paths + path].                                                 */
int mlmcb_packed_probe(int sampler, uint64_t seed, int l, int n_steps, uint64_t path_offset, uint64_t n_paths,
                       int device, void* cuda_stream, double* d_out);
/* The jump-stream records the Philox-mode Merton kernels use: d_out[3][n_records][n_paths] =
 * inter-arrival times (scaled by 1/lam; fp64 sampler: two per Philox block of the gap stream), jump-size
 * normals, Brownian normals of the sub-step ending at each jump (one block of the jump stream per record).  With mlmcb_sampler_probe this lets a test rebuild the exact draw streams of a Philox-mode
 * call and push them through the replay entry point.                                                */
int mlmcb_jump_probe(int sampler, uint64_t seed, int l, double lam, uint64_t path_offset, uint64_t n_paths,
                     uint64_t n_records, int device, void* cuda_stream, double* d_out);
/* Host-side Philox4x32-10 (same code the kernels use), for known-answer tests; no GPU needed. */
void mlmcb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Pipe micro-benchmarks used for the roofline denominators.  which: 0 DFMA, 1 DADD, 2 FFMA,
 * 3 IMAD.WIDE (mul.wide.u32+add), 4 LOP3, 5 MUFU.LG2, 6 I2F.U32, 7 F2F.F64.F32, 8 MUFU.SIN; two-instruction
 * mixes on independent chains: 9 DFMA+LOP3, 10 DFMA+IMAD.WIDE, 11 IMAD.WIDE+LOP3, 12 FFMA+LOP3, 13 DFMA+FFMA,
 * 14 DFMA+IMAD (the table behind "an INT32 instruction holds the dispatch port two cycles").
 * Returns warp-level instructions per second per GPU in *inst_per_s and the elapsed ms.      */
int mlmcb_pipe_probe(int which, int device, double* inst_per_s, float* ms);
/* 15 IMAD.HI, 16 DFMA+IMAD.HI, 17 IMAD.HI+LOP3, 18 IMAD.HI+IMAD complete the table (tools/pipe_probe.py).
 * mlmcb_mix_probe: the headline loop's per-step instruction mix (16 FP64, 6 IMAD.WIDE, 15 INT32) on independent
 * chains, `iters` passes on every SM.  mode 0: every warp issues the whole mix, finely interleaved (two warps per
 * SMSP); mode 1: warp-specialised (two FP64 warps and one Philox warp per SMSP).  One pass is two warp-steps per SMSP
 * in both modes; elapsed milliseconds in *ms (tools/mix_probe.py).                               */
int mlmcb_mix_probe(int mode, int device, int iters, float* ms);
/* Steps/sec the kernel's launch geometry: threads per block and resident blocks used.       */
int mlmcb_launch_geometry(int model, int payoff, int M, int flags, int device,
                          int* threads_per_block, int* max_blocks);

void mlmcb_params_fill_derived(mlmcb_params* p);   /* disc = exp(-r*T), J_bar = exp(m+s^2/2)-1 */
const char* mlmcb_last_error(void);
int mlmcb_version(void);                            /* 102 (101: flags == 0 meant the fp32 sampler; 100: mlmcb_params without amer_boundary/amer_levels) */

#ifdef __cplusplus
}
#endif
#endif /* MLMCB_H */
