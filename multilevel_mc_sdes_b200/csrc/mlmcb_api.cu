// mlmcb_api.cu - C ABI (include/mlmcb.h) over the kernels in mlmcb_kernels.cuh.
// Host side only: argument checking, the per-launch constant block (evaluated in the same
// order numpy evaluates the reference's expressions), kernel selection, launch geometry,
// stream-ordered workspace, and the pipe micro-benchmarks.  No torch types, no CPU fallback.
#include "../../include/mlmcb.h"
#include "mlmcb_pick.cuh"

#include <math.h>
#include <algorithm>
#include <map>
#include <stddef.h>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace mlmcb;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(MLMCB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DeviceInfo { int sms = 0; cudaMemPool_t pool = nullptr; };
DeviceInfo g_dev[64];
std::mutex g_dev_mu;               // first-use initialisation of g_dev / g_host entries

int device_info(int dev, DeviceInfo** out) {
    if (dev < 0 || dev >= 64) return fail(MLMCB_EINVAL, "device %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_dev_mu);
    DeviceInfo& d = g_dev[dev];
    if (!d.sms) {
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        // a pool of the library's own for the stream-ordered workspace: its release threshold keeps freed
        // blocks cached without changing the trimming behaviour of the process's default pool
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        CU(cudaMemPoolCreate(&d.pool, &props));
        unsigned long long keep = ~0ull;
        CU(cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        d.sms = sms;
    }
    *out = &d;
    return 0;
}

int ipow_checked(int M, int l, ull* out) {
    ull n = 1;
    for (int i = 0; i < l; ++i) {
        if (n > (1ull << 33) / (ull)M) return fail(MLMCB_EINVAL, "M**l = %d**%d exceeds 2^33 fine steps", M, l);
        n *= (ull)M;
    }
    *out = n;
    return 0;
}

// The per-launch constants.  Each reference-order scalar is one IEEE double operation per
// Python operator, in Python's evaluation order (citations: /root/reference/MLMC_RSCAM.py).
uint32_t run_blocks_for(ull nsteps);

int make_consts(const mlmcb_params* p, int model, int payoff, int l, int M, ull n_paths, ull seed,
                ull path_offset, LevelConsts* C) {
    if (!p) return fail(MLMCB_EINVAL, "params is NULL");
    if (model != MLMCB_GBM && model != MLMCB_MERTON) return fail(MLMCB_EINVAL, "unknown model %d", model);
    if (payoff < 0 || payoff > 4) return fail(MLMCB_EINVAL, "unknown payoff %d", payoff);
    if (payoff == MLMCB_AMERICAN) {
        if (model != MLMCB_GBM) return fail(MLMCB_ENOTIMPL, "the American put is implemented for GBM only (Amer_GBM)");
        if (M != 2) return fail(MLMCB_EINVAL, "M must be equal to 2 for American Option pricer. Please use M=2.");   // :1525
    }
    if (l < 0 || l > 255) return fail(MLMCB_EINVAL, "level %d out of range", l);
    if (M < 2) return fail(MLMCB_EINVAL, "coarseness factor M=%d must be >= 2", M);
    if (path_offset + n_paths >= (1ull << 48)) return fail(MLMCB_EINVAL, "path ids must stay below 2^48");
    if (!(p->T > 0) || !(p->sig >= 0) || !(p->X0 > 0)) return fail(MLMCB_EINVAL, "need T>0, sig>=0, X0>0");
    if (model == MLMCB_MERTON && !(p->lam >= 0)) return fail(MLMCB_EINVAL, "need lam>=0");
    memset(C, 0, sizeof *C);
    ull nsteps;
    if (int rc = ipow_checked(M, l, &nsteps)) return rc;
    C->n_paths = n_paths; C->path_offset = path_offset; C->nsteps = nsteps;
    C->M = M; C->l = l; C->payoff = payoff; C->is_l0 = (l == 0);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { C->k0[r] = k0; C->k1[r] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
    C->ctr3_tag = (uint32_t)l << 16;
    const double disc = p->disc > 0 ? p->disc : exp(-p->r * p->T);                   // :80
    C->X0 = p->X0; C->K = p->K; C->T = p->T; C->sig = p->sig;
    C->Md = (double)M;
    C->dt = p->T / (double)nsteps;                                                   // :334
    C->sqrt_dt = sqrt(C->dt);                                                        // :335
    C->Mdt = C->Md * C->dt;                                                          // M*dt :347
    if (model == MLMCB_GBM) {
        C->mu = p->r + p->drift;                                                     // :1347
    } else {
        const double J_bar = (p->J_bar == p->J_bar && p->J_bar != 0.0) ? p->J_bar
                             : exp(p->jumpmean + 0.5 * p->jumpstd * p->jumpstd) - 1; // :1291
        C->mu = p->r - p->lam * J_bar;                                               // :1314
        C->lam = p->lam; C->inv_lam = 1.0 / p->lam; C->jm = p->jumpmean; C->js = p->jumpstd;
    }
    C->disc = disc;
    C->discK = disc * p->K;                                                          // :131
    C->look_cf = 1 - p->beta * p->sig * sqrt(C->dt);                                 // :105
    C->look_cc = 1 - p->beta * p->sig * sqrt(C->Md * C->dt);                         // :110
    C->a_f = C->mu * C->dt; C->b_f = C->sig * C->sqrt_dt; C->a_c = C->mu * C->Mdt;
    C->avg_wf = C->dt / C->T; C->avg_wc = C->Mdt / C->T;
    C->poly_q4 = MLMCB_Q4; C->poly_c4 = MLMCB_C4;
    C->inv_blk_dt = 1.0 / (4.0 * C->dt);
    C->inv_sqrt_dt = 1.0 / C->sqrt_dt;
    C->inv_T = 1.0 / C->T;
    C->run_blocks = run_blocks_for(C->nsteps);
    // table-driven fp64 Box-Muller (tools/gen_f64bm_tables.py): G(x) = x + x^2 Q(x); sin(pi d)/d and (cos(pi d) - 1)/d^2 in d^2
    static const double lq[5] = {0x1.fffffffffeaabp-3, 0x1.555555555430cp-4, 0x1.00002aaab1aabp-5, 0x1.9999e2be38548p-7, 0.0};
    static const double ts[3] = {0x1.921fb54442d18p+1, -0x1.4abbce625a8c0p+2, 0x1.466ba9b398947p+1};
    static const double tc[2] = {-0x1.3bd3cc9be3b30p+2, 0x1.03c1db247a00ep+2};
    memcpy(C->lq, lq, sizeof lq); memcpy(C->ts, ts, sizeof ts); memcpy(C->tc, tc, sizeof tc);
    C->neg2ln2 = -0x1.62e42fefa39efp+0;
    C->one_hi = 0x3ff00000u;
    C->amer_vals = (payoff == MLMCB_AMERICAN && p->amer_boundary && l < p->amer_levels) ? p->amer_boundary[l] : nullptr;
    C->ln2_hi = 6.93147180369123816490e-01; C->ln2_lo = 1.90821492927058770002e-10;
    { double f = 1.0; for (int k = 2; k <= 13; ++k) { f *= (double)k; C->ex[k - 2] = 1.0 / f; } }
    C->r = p->r; C->boundary_c = p->boundary_c;
    C->sig2 = pow(p->sig, 2.0);                                                      // self.sig**2
    C->half_sig2 = 0.5 * C->sig2;
    C->c_mil = C->half_sig2 * C->dt;
    C->a_fm = (C->mu - C->half_sig2) * C->dt;
    C->a_cm = (C->mu - C->half_sig2) * C->Mdt;
    C->amer_cneg = -2.0 / (C->sig2 * C->dt);
    return 0;
}

int fn_of(int payoff) { return payoff == MLMCB_ASIAN ? FN_AVG : payoff == MLMCB_LOOKBACK ? FN_MIN : FN_FINAL; }
int mt_of(int M) { return M == 2 ? 2 : M == 4 ? 4 : 0; }


// Merton levels below this many fine steps run the event-driven kernel (mjd_flat_kernel), the rest the
// block-uniform one; MLMCB_MJD_FLAT_BELOW / MLMCB_MJD_RUN_BLOCKS in the environment override (tuning).
ull flat_below() {
    static const ull v = [] {
        const char* e = getenv("MLMCB_MJD_FLAT_BELOW");
        return e ? strtoull(e, nullptr, 10) : (ull)MLMCB_MJD_FLAT_BELOW;
    }();
    return v;
}
uint32_t run_blocks_for(ull nsteps) {
    static const long forced = [] {
        const char* e = getenv("MLMCB_MJD_RUN_BLOCKS");
        return e ? strtol(e, nullptr, 10) : 0L;
    }();
    if (forced > 0) return (uint32_t)forced;
    // a lane waits about half a run per event, an event phase costs about as much as 20 fine steps
    const double r = sqrt((double)nsteps * 20.0) / 4.0;
    return r < 4.0 ? 4u : r > 4096.0 ? 4096u : (uint32_t)r;
}

LevelKernel pick_kernel(int model, int payoff, int M, int flags, ull nsteps) {
    const int fn = fn_of(payoff), mt = mt_of(M);
    const bool f64 = (flags & MLMCB_SAMPLER_MASK) == MLMCB_SAMPLER_F64BM;
    if (model == MLMCB_GBM) {
        // the coarsest levels have kernels of their own: 1 or 2 fine steps, and 4 (M = 4, l = 1 / M = 2, l = 2)
        const int packed = nsteps <= 2 ? 1 : (nsteps == 4 && M == 4) ? 4 : (nsteps == 4 && M == 2) ? -4 : 0;
        if (packed) return f64 ? pick_gbm_packed_s<1>(fn, flags, packed) : pick_gbm_packed_s<0>(fn, flags, packed);
        // long paths with the default sampler: the software-pipelined build of the same loop
        if (f64 && mt != 0 && !(flags & (MLMCB_MILSTEIN | MLMCB_FP32_PATHS)) && nsteps >= (ull)MLMCB_PIPE_FROM && (nsteps & 7) == 0)
            return pick_gbm_long_s<1>(fn, mt);
        return f64 ? pick_gbm_s<1>(fn, mt, flags) : pick_gbm_s<0>(fn, mt, flags);
    }
    static const char* force = getenv("MLMCB_MJD_WALK");                 // "batch" | "event" | "legacy": tuning and the walk-agreement test
    if (mjd_use_factor(flags, mt, nsteps) && !force) return f64 ? pick_merton_factor_s<1>(fn, mt) : pick_merton_factor_s<0>(fn, mt);
    if (nsteps < flat_below()) {
        const bool batch = force ? (force[0] == 'b' && nsteps > 1) : mjd_use_batch(f64 ? SAMP_F64BM : SAMP_F32BM, nsteps);
        return f64 ? pick_merton_flat_s<1>(fn, mt, flags, batch) : pick_merton_flat_s<0>(fn, mt, flags, batch);
    }
    return f64 ? pick_merton_uniform_s<1>(fn, mt, flags) : pick_merton_uniform_s<0>(fn, mt, flags);
}

// Launch geometry: a grid-stride loop over paths.  The grid never exceeds the number of
// co-resident blocks (148 SMs x occupancy), and is shrunk so every thread gets the same
// number of paths (no ragged tail wave).
template <typename K>
int geometry(K kernel, int dev, ull n_paths, int* grid, int* max_blocks_out) {
    DeviceInfo* d;
    if (int rc = device_info(dev, &d)) return rc;
    int per_sm = 0;
    {
        // per (device, kernel), once: kernels that park jump records in shared memory ask for the
        // large carve-out, so the occupancy computed here is the one the launch gets
        static std::mutex mu;
        static std::map<std::pair<int, const void*>, int> seen;
        std::lock_guard<std::mutex> lock(mu);
        auto it = seen.find({dev, (const void*)kernel});
        if (it == seen.end()) {
            cudaFuncAttributes fa;
            CU(cudaFuncGetAttributes(&fa, kernel));
            if (fa.sharedSizeBytes > 4096)
                CU(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
            if (per_sm < 1) per_sm = 1;
            seen[{dev, (const void*)kernel}] = per_sm;
        } else {
            per_sm = it->second;
        }
    }
    const ull max_threads = (ull)d->sms * per_sm * kThreads;
    const ull per_thread = (n_paths + max_threads - 1) / max_threads;
    const ull threads = per_thread ? (n_paths + per_thread - 1) / per_thread : 1;
    ull g = (threads + kThreads - 1) / kThreads;
    if (g < 1) g = 1;
    *grid = (int)g;
    if (max_blocks_out) *max_blocks_out = d->sms * per_sm;
    return 0;
}

struct Workspace {
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    void* base = nullptr;
};

int pool_alloc(void** ptr, size_t bytes, int device, cudaStream_t st) {
    DeviceInfo* d;
    if (int rc = device_info(device, &d)) return rc;
    CU(cudaMallocFromPoolAsync(ptr, bytes, d->pool, st));
    return 0;
}

int ws_alloc(int grid, int device, cudaStream_t st, Workspace* w) {
    const size_t bytes = (size_t)grid * 8 * sizeof(double) + 256;
    if (int rc = pool_alloc(&w->base, bytes, device, st)) return rc;
    w->ticket = (unsigned*)w->base;
    w->partials = (double*)((char*)w->base + 256);
    CU(cudaMemsetAsync(w->ticket, 0, sizeof(unsigned), st));
    return 0;
}
int ws_free(Workspace* w, cudaStream_t st) {
    CU(cudaFreeAsync(w->base, st));
    return 0;
}

int check_flags(int flags) {
    const int samp = flags & MLMCB_SAMPLER_MASK;
    if (samp != MLMCB_SAMPLER_F32BM && samp != MLMCB_SAMPLER_F64BM) return fail(MLMCB_EINVAL, "unknown sampler %d", samp);
    if (flags & ~(MLMCB_SAMPLER_MASK | MLMCB_FP32_PATHS | MLMCB_ANTITHETIC | MLMCB_MILSTEIN))
        return fail(MLMCB_EINVAL, "unknown flag bits 0x%x", flags);
    // every sampler x scheme x estimator combination is built for fp64 paths; the fp32 opt-in covers Euler-Maruyama
    if ((flags & MLMCB_MILSTEIN) && (flags & MLMCB_FP32_PATHS))
        return fail(MLMCB_ENOTIMPL, "Milstein is built for fp64 path arithmetic only");
    return 0;
}

// looper(anti=True): only Euro_GBM / Asian_GBM define anti_payoff (:1379/:1406); M must be even (:178)
int check_anti(int model, int payoff, int M) {
    if (model != MLMCB_GBM || (payoff != MLMCB_EURO && payoff != MLMCB_ASIAN))
        return fail(MLMCB_ENOTIMPL, "Option instance has no implemented anti_payoff method");
    if (M % 2 != 0)
        return fail(MLMCB_EINVAL, "Cannot calculate antithetic estimator for odd coarseness factor M - please use even M");
    return 0;
}

LevelKernel pick_anti_kernel(int payoff, int M, int flags) {
    const int fn = fn_of(payoff), mt = mt_of(M);
    return (flags & MLMCB_SAMPLER_MASK) == MLMCB_SAMPLER_F64BM ? pick_gbm_anti_s<1>(fn, mt, flags) : pick_gbm_anti_s<0>(fn, mt, flags);
}

// American put: tabulate the per-step scalars, then run the path kernel (Philox or replay).
constexpr ull kAmerMaxSteps = 1ull << 22;

// host staging of a tabulated boundary: freed by the stream once the copy that reads it has run
void CUDART_CB free_host_rows(void* p) { free(p); }

int amer_table(const LevelConsts& C, int device, cudaStream_t st, AmerRow** tab) {
    if (C.nsteps > kAmerMaxSteps) return fail(MLMCB_EINVAL, "American put: M**l above 2^22 fine steps is not supported");
    *tab = nullptr;
    if (int rc = pool_alloc((void**)tab, (size_t)(C.nsteps + 1) * sizeof(AmerRow), device, st)) return rc;
    if (C.amer_vals) {
        // tabulated boundary: the same rows as amer_table_kernel, built here from the caller's values
        // (b on the grid | at the fine mid-points | at the coarse mid-points), same operation order
        const ull n = C.nsteps;
        const double *bg = C.amer_vals, *bf = bg + n + 1, *bc = bf + n;
        AmerRow* rows = (AmerRow*)malloc((n + 1) * sizeof(AmerRow));
        if (!rows) { cudaFreeAsync(*tab, st); *tab = nullptr; return fail(MLMCB_ECUDA, "out of host memory for the boundary table"); }
        for (ull j = 0; j <= n; ++j) {
            AmerRow row = {0, 0, 0, 0, 0, 0, 0, 0};
            row.bR = bg[j];
            if (j >= 1) {
                const double t_ = (double)(j - 1) * C.dt, tm = t_ + C.dt / 2.0;                 // :1546, :1562
                row.mkF = fmax(C.K - bf[j - 1], 0.0);
                row.eF = exp(-C.r * tm);                                                       // :1571
                if (j % 2 == 0) {
                    const double tc = (double)(j - 2) * C.dt, tcm = tc + C.Mdt / 2.0;           // :1580, :1590
                    row.mkC = fmax(C.K - bc[j / 2 - 1], 0.0);
                    row.eC = exp(-C.r * tcm);                                                  // :1605
                    row.mb1 = 0.5 * (row.bR + bg[j - 2]);                                      // :1599
                }
            }
            rows[j] = row;
        }
        // pageable source: the copy is staged by the runtime before the call returns for small tables and
        // otherwise ordered on the stream; the staging block is released by a host callback AFTER the copy,
        // so the entry point stays asynchronous (no cudaStreamSynchronize)
        cudaError_t e = cudaMemcpyAsync(*tab, rows, (n + 1) * sizeof(AmerRow), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaLaunchHostFunc(st, free_host_rows, rows);
        if (e != cudaSuccess) {
            cudaStreamSynchronize(st);
            free(rows);
            cudaFreeAsync(*tab, st);
            *tab = nullptr;
            return fail(MLMCB_ECUDA, "boundary table upload failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    const int threads = 128;
    const int blocks = (int)((C.nsteps + 1 + threads - 1) / threads);
    amer_table_kernel<<<blocks, threads, 0, st>>>(C, *tab);
    if (cudaError_t e = cudaGetLastError()) {
        cudaFreeAsync(*tab, st);
        *tab = nullptr;
        return fail(MLMCB_ECUDA, "amer_table_kernel launch failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

int amer_level_sums(const LevelConsts& C, int flags, int device, cudaStream_t st, double* d_out7, double* d_Pf,
                    double* d_Pc) {
    if (flags & (MLMCB_FP32_PATHS | MLMCB_ANTITHETIC | MLMCB_MILSTEIN))
        return fail(MLMCB_ENOTIMPL, "the American put is built for fp64 Euler-Maruyama paths without the antithetic estimator");
    typedef void (*K)(const LevelConsts, const AmerRow*, double*, unsigned*, double*, double*, double*);
    K k = (flags & MLMCB_SAMPLER_MASK) == MLMCB_SAMPLER_F64BM ? amer_level_kernel<SAMP_F64BM> : amer_level_kernel<SAMP_F32BM>;
    int grid = 1;
    if (int rc = geometry(k, device, C.n_paths, &grid, nullptr)) return rc;
    AmerRow* tab = nullptr;
    if (int rc = amer_table(C, device, st, &tab)) return rc;
    Workspace w;
    if (int rc = ws_alloc(grid, device, st, &w)) { cudaFreeAsync(tab, st); return rc; }
    k<<<grid, kThreads, 0, st>>>(C, tab, w.partials, w.ticket, d_out7, d_Pf, d_Pc);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tab, st);
    if (int rc = ws_free(&w, st)) return rc;
    if (e != cudaSuccess) return fail(MLMCB_ECUDA, "amer_level_kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int level_sums_impl(const mlmcb_params* p, int model, int payoff, int l, int M, ull n_paths, ull seed,
                    ull path_offset, int flags, int device, cudaStream_t st, double* d_out7, double* d_Pf,
                    double* d_Pc, double* d_raw = nullptr) {
    if (!d_out7) return fail(MLMCB_EINVAL, "d_out7 is NULL");
    if (int rc = check_flags(flags)) return rc;
    LevelConsts C;
    if (int rc = make_consts(p, model, payoff, l, M, n_paths, seed, path_offset, &C)) return rc;
    C.raw = d_raw;
    if (flags & MLMCB_ANTITHETIC) {
        if (int rc = check_anti(model, payoff, M)) return rc;
    }
    if (n_paths == 0) { CU(cudaMemsetAsync(d_out7, 0, 7 * sizeof(double), st)); return 0; }
    if (payoff == MLMCB_AMERICAN) return amer_level_sums(C, flags, device, st, d_out7, d_Pf, d_Pc);
    // l = 0 has no antithetic twin (diffusion_anti_path :212-214 takes one plain step and anti_EA_payoff :310
    // returns the plain Pf), so the sums come from the plain kernel and its packed draws; the antithetic
    // kernel still serves per-path outputs there (its coarse output is X0, :214)
    const bool anti = (flags & MLMCB_ANTITHETIC) && !(C.is_l0 && !d_Pf && !d_Pc && !d_raw);
    LevelKernel k = anti ? pick_anti_kernel(payoff, M, flags)
                         : pick_kernel(model, payoff, M, flags & ~MLMCB_ANTITHETIC, C.nsteps);
    int grid = 1;
    if (int rc = geometry(k, device, n_paths, &grid, nullptr)) return rc;
    Workspace w;
    if (int rc = ws_alloc(grid, device, st, &w)) return rc;
    k<<<grid, kThreads, 0, st>>>(C, w.partials, w.ticket, d_out7, d_Pf, d_Pc);
    CU(cudaGetLastError());
    return ws_free(&w, st);
}

// pinned staging + side streams for the _host entry points (per device, created lazily)
struct HostSide {
    double* pinned = nullptr;      // [kMaxSweep][8]
    double* dev = nullptr;         // same shape on the device
    cudaStream_t side[8] = {};
    cudaEvent_t fork = nullptr, join[8] = {};
    // persistent workspace of the one-launch sweep: partial rows, the ticket (the kernel leaves it zero) and an
    // event that orders consecutive sweeps issued on different streams
    double* sweep_partials = nullptr;
    unsigned* sweep_ticket = nullptr;
    cudaEvent_t sweep_done = nullptr;
    bool sweep_pending = false;
    bool ready = false;
};
constexpr int kMaxSweep = 64;
constexpr int kSweepRows = 8192;   // upper bound on the grid of a sweep launch
HostSide g_host[64];
std::mutex g_host_mu[64];          // the staging buffers of a device serve one _host call at a time

int host_side(int dev, HostSide** out) {
    HostSide& h = g_host[dev];
    if (!h.ready) {
        CU(cudaHostAlloc((void**)&h.pinned, kMaxSweep * 8 * sizeof(double), cudaHostAllocDefault));
        CU(cudaMalloc((void**)&h.dev, kMaxSweep * 8 * sizeof(double)));
        CU(cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming));
        CU(cudaMalloc((void**)&h.sweep_partials, (size_t)kSweepRows * 8 * sizeof(double)));
        CU(cudaMalloc((void**)&h.sweep_ticket, 256));
        CU(cudaMemset(h.sweep_ticket, 0, 256));
        CU(cudaEventCreateWithFlags(&h.sweep_done, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) {
            CU(cudaStreamCreateWithFlags(&h.side[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h.join[i], cudaEventDisableTiming));
        }
        h.ready = true;
    }
    *out = &h;
    return 0;
}

// ---- one launch per sweep ----------------------------------------------------------------------
// Relative cost of one path in SMSP cycles per warp (profiles/r0*_gbm_levels.txt, r0*_merton_levels.txt): only the
// proportions matter - they decide how many blocks of the launch each level gets; the block scheduler evens
// out what the model gets wrong because the grid is several waves deep.
double path_cost(int model, bool f64, ull nsteps, double lamT, bool factor) {
    const double step = f64 ? 80.0 : 38.0, n = (double)nsteps;
    if (model == MLMCB_GBM) {
        if (nsteps == 1) return f64 ? 260.0 : 130.0;
        if (nsteps == 2) return f64 ? 520.0 : 265.0;
        return (f64 ? 400.0 : 230.0) + step * n;
    }
    const double event = (f64 ? 1600.0 : 800.0) * (lamT + 1.0);
    if (nsteps == 1) return event;
    // factor walk: arrivals + records per path, then steps whose extra cost falls with the share of jump-free octets
    if (factor) return (f64 ? 1000.0 : 850.0) * (lamT + 1.0) + step * n * (1.04 + 0.7 * lamT * exp(-n / 100.0));
    if (nsteps < flat_below()) return 1.3 * event + 2.5 * step * n;
    return 6.0 * event + 1.05 * step * n;
}

// Tries to run the sweep as ONE kernel; *done = false (and nothing launched) if this request is not covered
// (antithetic, fp32 paths, American put, more than kSweepMax non-empty levels).  d_out: [n_levels][8] on the device.
int sweep_fused(const mlmcb_params* p, int model, int payoff, int M, int n_levels, const int* levels,
                const uint64_t* n_paths, const uint64_t* path_offsets, uint64_t seed, int flags, int device,
                cudaStream_t st, HostSide* h, double* d_out, bool* done) {
    *done = false;
    static const bool off = getenv("MLMCB_NO_FUSED_SWEEP") != nullptr;      // A/B timing and the agreement test
    if (off || payoff == MLMCB_AMERICAN || (flags & (MLMCB_ANTITHETIC | MLMCB_FP32_PATHS))) return 0;
    int live = 0;
    for (int i = 0; i < n_levels; ++i) live += n_paths[i] != 0;
    if (live == 0 || live > kSweepMax) return 0;
    if (int rc = check_flags(flags)) return rc;
    const bool f64 = (flags & MLMCB_SAMPLER_MASK) == MLMCB_SAMPLER_F64BM;
    const int fn = fn_of(payoff), mt = mt_of(M);
    SweepKernel k = model == MLMCB_GBM ? (f64 ? pick_gbm_sweep_s<1>(fn, mt, flags) : pick_gbm_sweep_s<0>(fn, mt, flags))
                                       : (f64 ? pick_merton_sweep_s<1>(fn, mt, flags) : pick_merton_sweep_s<0>(fn, mt, flags));
    static thread_local SweepConsts S;            // 25 KB: too big for comfort on the stack
    memset(&S, 0, offsetof(SweepConsts, lv));
    S.n_levels = live;
    S.n_rows = n_levels;
    // deepest level first (longest serial path); ties keep the caller's order
    std::vector<int> order;
    for (int i = 0; i < n_levels; ++i) if (n_paths[i]) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return levels[a] > levels[b]; });
    double cost[kSweepMax], total = 0.0, total_long = 0.0;
    ull cap[kSweepMax];
    for (int j = 0; j < live; ++j) {
        const int i = order[j];
        LevelConsts* C = &S.lv[j];
        if (int rc = make_consts(p, model, payoff, levels[i], M, n_paths[i], seed, path_offsets[i], C)) return rc;
        S.row[j] = i;
        ull per_thread = 1;
        if (model == MLMCB_GBM) {
            if (C->nsteps <= 2) { S.kind[j] = SK_PACKED; per_thread = C->nsteps == 1 ? 4 : 2; }
            else if (C->nsteps == 4 && mt != 0) S.kind[j] = SK_FOUR;
            else S.kind[j] = SK_LONG;
        } else {
            static const bool legacy = getenv("MLMCB_MJD_WALK") != nullptr;
            S.kind[j] = (mjd_use_factor(flags, mt, C->nsteps) && !legacy) ? SK_MJD_FACTOR
                        : C->nsteps < flat_below() ? SK_MJD_FLAT : SK_MJD_UNIFORM;
        }
        cost[j] = (double)n_paths[i] * path_cost(model, f64, C->nsteps, model == MLMCB_MERTON ? p->lam * p->T : 0.0,
                                                 S.kind[j] == SK_MJD_FACTOR);
        total += cost[j];
        if (C->nsteps >= 64) total_long += cost[j];
        const ull threads = (n_paths[i] + per_thread - 1) / per_thread + (per_thread > 1 ? 1 : 0);
        cap[j] = (threads + kThreads - 1) / kThreads;                   // more blocks than this would idle
    }
    // One launch pays where launch latency matters: sweeps of up to a few milliseconds of work (the iterations of a
    // latency-bound mlmc run).  A long sweep - a fixed-level sweep of Giles_plot (:1775-1777), the iterations of a
    // tight-eps run - is better served by each level's own kernel, which is built for that level's path length
    // (four octets in flight for long GBM paths, its own occupancy for each Merton walk) where the sweep kernel is one
    // build for all.  MLMCB_FUSED_SWEEP_MS moves the threshold (tuning).
    static const double max_ms = [] { const char* e = getenv("MLMCB_FUSED_SWEEP_MS"); return e ? atof(e) : 3.0; }();
    (void)total_long;
    if (total / 32.0 > max_ms * 1e-3 * 148 * 4 * 1.9e9) return 0;
    int grid0 = 0, resident = 0;
    if (int rc = geometry(k, device, ~0ull >> 16, &grid0, &resident)) return rc;
    const double budget = (double)(resident * 4 < kSweepRows - kSweepMax ? resident * 4 : kSweepRows - kSweepMax);   // four waves
    unsigned first = 0;
    for (int j = 0; j < live; ++j) {
        ull b = (ull)(budget * cost[j] / total + 0.5);
        if (b < 1) b = 1;
        if (b > cap[j]) b = cap[j];
        S.first[j] = first;
        first += (unsigned)b;
    }
    S.first[live] = first;
    if (h->sweep_pending) CU(cudaStreamWaitEvent(st, h->sweep_done, 0));     // the workspace serves one sweep at a time
    k<<<first, kThreads, 0, st>>>(S, h->sweep_partials, h->sweep_ticket, d_out);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->sweep_done, st));
    h->sweep_pending = true;
    *done = true;
    return 0;
}

// the sweep as concurrent per-level launches on the side streams, rows of 8 doubles into d_out
int sweep_per_level(const mlmcb_params* p, int model, int payoff, int M, int n_levels, const int* levels,
                    const uint64_t* n_paths, const uint64_t* path_offsets, uint64_t seed, int flags, int device,
                    cudaStream_t st, HostSide* h, double* d_out) {
    // fork: side streams wait for whatever precedes on the caller's stream (and the zero-fill)
    CU(cudaMemsetAsync(d_out, 0, (size_t)n_levels * 8 * sizeof(double), st));
    CU(cudaEventRecord(h->fork, st));
    bool used[8] = {};
    // deepest level first: it is the longest-running launch of the sweep
    for (int i = n_levels - 1, slot = 0; i >= 0; --i) {
        if (n_paths[i] == 0) continue;
        cudaStream_t s = h->side[slot & 7];
        if (!used[slot & 7]) { CU(cudaStreamWaitEvent(s, h->fork, 0)); used[slot & 7] = true; }
        if (int rc = level_sums_impl(p, model, payoff, levels[i], M, n_paths[i], seed, path_offsets[i], flags,
                                     device, s, d_out + (size_t)i * 8, nullptr, nullptr)) {
            cudaDeviceSynchronize();      // launches already issued on side streams must not outlive the call
            return rc;
        }
        ++slot;
    }
    for (int s = 0; s < 8; ++s) {
        if (!used[s]) continue;
        CU(cudaEventRecord(h->join[s], h->side[s]));
        CU(cudaStreamWaitEvent(st, h->join[s], 0));
    }
    return 0;
}

int sweep_impl(const mlmcb_params* p, int model, int payoff, int M, int n_levels, const int* levels,
               const uint64_t* n_paths, const uint64_t* path_offsets, uint64_t seed, int flags, int device,
               cudaStream_t st, HostSide* h, double* d_out) {
    if (n_levels == 0) return 0;
    bool done = false;
    int live = 0;
    for (int i = 0; i < n_levels; ++i) live += n_paths[i] != 0;
    // a single level keeps its own kernel (one wave, equal paths per thread); two or more share one launch
    if (live >= 2)
    if (int rc = sweep_fused(p, model, payoff, M, n_levels, levels, n_paths, path_offsets, seed, flags, device, st, h,
                             d_out, &done))
        return rc;
    if (done) return 0;
    return sweep_per_level(p, model, payoff, M, n_levels, levels, n_paths, path_offsets, seed, flags, device, st, h, d_out);
}

}  // namespace

extern "C" {

const char* mlmcb_last_error(void) { return g_err; }
int mlmcb_version(void) { return 102; }      // 102: flags == 0 selects the fp64 sampler (fp32 is the opt-in); 101: mlmcb_params grew amer_boundary / amer_levels

void mlmcb_params_fill_derived(mlmcb_params* p) {
    if (!p) return;
    p->disc = exp(-p->r * p->T);
    p->J_bar = exp(p->jumpmean + 0.5 * p->jumpstd * p->jumpstd) - 1;
}

void mlmcb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) { c = philox_round(c, k0, k1); k0 += PHILOX_W0; k1 += PHILOX_W1; }
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}

int mlmcb_level_sums(const mlmcb_params* p, int model, int payoff, int l, int M, uint64_t n_paths,
                     uint64_t seed, uint64_t path_offset, int flags, int device, void* cuda_stream,
                     double* d_out7, double* d_Pf, double* d_Pc) {
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    return level_sums_impl(p, model, payoff, l, M, n_paths, seed, path_offset, flags, device,
                           (cudaStream_t)cuda_stream, d_out7, d_Pf, d_Pc);
}

int mlmcb_path_rows(int payoff, int flags) {
    if (payoff == MLMCB_AMERICAN) return 6;
    return (flags & MLMCB_ANTITHETIC) ? 3 : 4;
}

int mlmcb_level_paths(const mlmcb_params* p, int model, int payoff, int l, int M, uint64_t n_paths,
                      uint64_t seed, uint64_t path_offset, int flags, int device, void* cuda_stream,
                      double* d_out7, double* d_raw) {
    if (!d_raw) return fail(MLMCB_EINVAL, "d_raw is NULL");
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    return level_sums_impl(p, model, payoff, l, M, n_paths, seed, path_offset, flags, device,
                           (cudaStream_t)cuda_stream, d_out7, nullptr, nullptr, d_raw);
}

int mlmcb_level_sums_host(const mlmcb_params* p, int model, int payoff, int l, int M, uint64_t n_paths,
                          uint64_t seed, uint64_t path_offset, int flags, int device, void* cuda_stream,
                          double* h_out7) {
    if (!h_out7) return fail(MLMCB_EINVAL, "h_out7 is NULL");
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    if (device < 0 || device >= 64) return fail(MLMCB_EINVAL, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_host_mu[device]);
    HostSide* h;
    if (int rc = host_side(device, &h)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (int rc = level_sums_impl(p, model, payoff, l, M, n_paths, seed, path_offset, flags, device, st, h->dev,
                                 nullptr, nullptr))
        return rc;
    CU(cudaMemcpyAsync(h->pinned, h->dev, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(h_out7, h->pinned, 7 * sizeof(double));
    return 0;
}

int mlmcb_level_sweep_host(const mlmcb_params* p, int model, int payoff, int M, int n_levels,
                           const int* levels, const uint64_t* n_paths, const uint64_t* path_offsets,
                           uint64_t seed, int flags, int device, void* cuda_stream, double* h_out) {
    if (!levels || !n_paths || !path_offsets || !h_out) return fail(MLMCB_EINVAL, "NULL array argument");
    if (n_levels < 0 || n_levels > kMaxSweep) return fail(MLMCB_EINVAL, "n_levels=%d out of range (max %d)", n_levels, kMaxSweep);
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    if (device < 0 || device >= 64) return fail(MLMCB_EINVAL, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_host_mu[device]);
    HostSide* h;
    if (int rc = host_side(device, &h)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (int rc = sweep_impl(p, model, payoff, M, n_levels, levels, n_paths, path_offsets, seed, flags, device, st, h, h->dev))
        return rc;
    CU(cudaMemcpyAsync(h->pinned, h->dev, (size_t)n_levels * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < n_levels; ++i) memcpy(h_out + (size_t)i * 7, h->pinned + (size_t)i * 8, 7 * sizeof(double));
    return 0;
}

int mlmcb_level_sweep(const mlmcb_params* p, int model, int payoff, int M, int n_levels,
                      const int* levels, const uint64_t* n_paths, const uint64_t* path_offsets,
                      uint64_t seed, int flags, int device, void* cuda_stream, double* d_out) {
    if (!levels || !n_paths || !path_offsets || !d_out) return fail(MLMCB_EINVAL, "NULL array argument");
    if (n_levels < 0 || n_levels > kMaxSweep) return fail(MLMCB_EINVAL, "n_levels=%d out of range (max %d)", n_levels, kMaxSweep);
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    if (device < 0 || device >= 64) return fail(MLMCB_EINVAL, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_host_mu[device]);
    HostSide* h;
    if (int rc = host_side(device, &h)) return rc;
    return sweep_impl(p, model, payoff, M, n_levels, levels, n_paths, path_offsets, seed, flags, device,
                      (cudaStream_t)cuda_stream, h, d_out);
}

int mlmcb_replay_gbm(const mlmcb_params* p, int payoff, int l, int M, uint64_t n_paths, const double* d_Z,
                     int arith, int device, void* cuda_stream, double* d_Pf, double* d_Pc, double* d_out7) {
    if (!d_Z || !d_Pf || !d_Pc) return fail(MLMCB_EINVAL, "NULL device pointer");
    const bool anti = (arith & MLMCB_ANTITHETIC) != 0, mil = (arith & MLMCB_MILSTEIN) != 0;
    arith &= ~(MLMCB_ANTITHETIC | MLMCB_MILSTEIN);
    if (arith != MLMCB_ARITH_FAST && arith != MLMCB_ARITH_EXACT) return fail(MLMCB_EINVAL, "unknown arith %d", arith);
    if (anti) {
        if (int rc = check_anti(MLMCB_GBM, payoff, M)) return rc;
    }
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LevelConsts C;
    if (int rc = make_consts(p, MLMCB_GBM, payoff, l, M, n_paths, 0, 0, &C)) return rc;
    if (n_paths == 0) return 0;
    if (payoff == MLMCB_AMERICAN) {
        if (anti || mil) return fail(MLMCB_ENOTIMPL, "the American put is built for Euler-Maruyama without the antithetic estimator");
        AmerRow* tab = nullptr;
        if (int rc = amer_table(C, device, st, &tab)) return rc;
        int grid = (int)((n_paths + kThreads - 1) / kThreads);
        if (grid > 65535) grid = 65535;
        Workspace w;
        if (int rc = ws_alloc(grid + 1, device, st, &w)) { cudaFreeAsync(tab, st); return rc; }
        double* out = d_out7 ? d_out7 : w.partials;
        if (arith == MLMCB_ARITH_EXACT)
            replay_amer_kernel<AR_EXACT><<<grid, kThreads, 0, st>>>(C, tab, d_Z, w.partials + 8, w.ticket, out, d_Pf, d_Pc);
        else
            replay_amer_kernel<AR_FAST><<<grid, kThreads, 0, st>>>(C, tab, d_Z, w.partials + 8, w.ticket, out, d_Pf, d_Pc);
        CU(cudaGetLastError());
        CU(cudaFreeAsync(tab, st));
        return ws_free(&w, st);
    }
    typedef void (*K)(const LevelConsts, const double*, double*, unsigned*, double*, double*, double*);
    static const K table[3][2] = {
        {replay_gbm_kernel<FN_FINAL, AR_FAST, 0>, replay_gbm_kernel<FN_FINAL, AR_EXACT, 0>},
        {replay_gbm_kernel<FN_AVG, AR_FAST, 0>, replay_gbm_kernel<FN_AVG, AR_EXACT, 0>},
        {replay_gbm_kernel<FN_MIN, AR_FAST, 0>, replay_gbm_kernel<FN_MIN, AR_EXACT, 0>}};
    static const K mil_table[3][2] = {
        {replay_gbm_kernel<FN_FINAL, AR_FAST, 1>, replay_gbm_kernel<FN_FINAL, AR_EXACT, 1>},
        {replay_gbm_kernel<FN_AVG, AR_FAST, 1>, replay_gbm_kernel<FN_AVG, AR_EXACT, 1>},
        {replay_gbm_kernel<FN_MIN, AR_FAST, 1>, replay_gbm_kernel<FN_MIN, AR_EXACT, 1>}};
    static const K anti_table[2][2][2] = {
        {{replay_gbm_anti_kernel<FN_FINAL, AR_FAST, 0>, replay_gbm_anti_kernel<FN_FINAL, AR_EXACT, 0>},
         {replay_gbm_anti_kernel<FN_AVG, AR_FAST, 0>, replay_gbm_anti_kernel<FN_AVG, AR_EXACT, 0>}},
        {{replay_gbm_anti_kernel<FN_FINAL, AR_FAST, 1>, replay_gbm_anti_kernel<FN_FINAL, AR_EXACT, 1>},
         {replay_gbm_anti_kernel<FN_AVG, AR_FAST, 1>, replay_gbm_anti_kernel<FN_AVG, AR_EXACT, 1>}}};
    K k = anti ? anti_table[mil ? 1 : 0][fn_of(payoff)][arith] : mil ? mil_table[fn_of(payoff)][arith] : table[fn_of(payoff)][arith];
    int grid = (int)((n_paths + kThreads - 1) / kThreads);
    if (grid > 65535) grid = 65535;
    Workspace w;
    if (int rc = ws_alloc(grid + 1, device, st, &w)) return rc;
    double* out = d_out7 ? d_out7 : w.partials;   // sums are always formed; row 0 is scratch if unwanted
    k<<<grid, kThreads, 0, st>>>(C, d_Z, w.partials + 8, w.ticket, out, d_Pf, d_Pc);
    CU(cudaGetLastError());
    return ws_free(&w, st);
}

int mlmcb_replay_merton(const mlmcb_params* p, int payoff, int l, int M, uint64_t n_paths,
                        const double* d_zW, const int64_t* d_offW, const double* d_zQ, const int64_t* d_offQ,
                        const double* d_E, const int64_t* d_offE, int arith, int device, void* cuda_stream,
                        double* d_Pf, double* d_Pc, double* d_out7) {
    if (!d_zW || !d_offW || !d_zQ || !d_offQ || !d_E || !d_offE || !d_Pf || !d_Pc)
        return fail(MLMCB_EINVAL, "NULL device pointer");
    const bool mil = (arith & MLMCB_MILSTEIN) != 0;
    arith &= ~MLMCB_MILSTEIN;
    if (arith != MLMCB_ARITH_FAST && arith != MLMCB_ARITH_EXACT) return fail(MLMCB_EINVAL, "unknown arith %d", arith);
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LevelConsts C;
    if (int rc = make_consts(p, MLMCB_MERTON, payoff, l, M, n_paths, 0, 0, &C)) return rc;
    if (n_paths == 0) return 0;
    typedef void (*K)(const LevelConsts, const double*, const long long*, const double*, const long long*,
                      const double*, const long long*, double*, unsigned*, double*, double*, double*);
    static const K table[3][2] = {
        {replay_mjd_kernel<FN_FINAL, AR_FAST, 0>, replay_mjd_kernel<FN_FINAL, AR_EXACT, 0>},
        {replay_mjd_kernel<FN_AVG, AR_FAST, 0>, replay_mjd_kernel<FN_AVG, AR_EXACT, 0>},
        {replay_mjd_kernel<FN_MIN, AR_FAST, 0>, replay_mjd_kernel<FN_MIN, AR_EXACT, 0>}};
    static const K mil_table[3][2] = {
        {replay_mjd_kernel<FN_FINAL, AR_FAST, 1>, replay_mjd_kernel<FN_FINAL, AR_EXACT, 1>},
        {replay_mjd_kernel<FN_AVG, AR_FAST, 1>, replay_mjd_kernel<FN_AVG, AR_EXACT, 1>},
        {replay_mjd_kernel<FN_MIN, AR_FAST, 1>, replay_mjd_kernel<FN_MIN, AR_EXACT, 1>}};
    K k = mil ? mil_table[fn_of(payoff)][arith] : table[fn_of(payoff)][arith];
    int grid = (int)((n_paths + kThreads - 1) / kThreads);
    if (grid > 65535) grid = 65535;
    Workspace w;
    if (int rc = ws_alloc(grid + 1, device, st, &w)) return rc;
    double* out = d_out7 ? d_out7 : w.partials;
    k<<<grid, kThreads, 0, st>>>(C, d_zW, (const long long*)d_offW, d_zQ, (const long long*)d_offQ, d_E,
                                 (const long long*)d_offE, w.partials + 8, w.ticket, out, d_Pf, d_Pc);
    CU(cudaGetLastError());
    return ws_free(&w, st);
}

int mlmcb_sampler_probe(int sampler, uint64_t seed, int l, uint64_t path_offset, uint64_t n_paths,
                        uint64_t n_steps, int device, void* cuda_stream, double* d_out) {
    if (!d_out) return fail(MLMCB_EINVAL, "d_out is NULL");
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    mlmcb_params p;
    memset(&p, 0, sizeof p);
    p.X0 = 1; p.K = 1; p.T = 1; p.sig = 1;
    LevelConsts C;
    if (int rc = make_consts(&p, MLMCB_GBM, MLMCB_EURO, 0, 2, n_paths, seed, path_offset, &C)) return rc;
    C.ctr3_tag = (uint32_t)l << 16;
    if (n_paths == 0 || n_steps == 0) return 0;
    const int grid = (int)((n_paths + kThreads - 1) / kThreads);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (sampler == MLMCB_SAMPLER_F32BM) sampler_probe_kernel<SAMP_F32BM><<<grid, kThreads, 0, st>>>(C, n_steps, d_out);
    else if (sampler == MLMCB_SAMPLER_F64BM) sampler_probe_kernel<SAMP_F64BM><<<grid, kThreads, 0, st>>>(C, n_steps, d_out);
    else return fail(MLMCB_EINVAL, "unknown sampler %d", sampler);
    CU(cudaGetLastError());
    return 0;
}

int mlmcb_packed_probe(int sampler, uint64_t seed, int l, int n_steps, uint64_t path_offset, uint64_t n_paths,
                       int device, void* cuda_stream, double* d_out) {
    if (!d_out) return fail(MLMCB_EINVAL, "d_out is NULL");
    if (n_steps != 1 && n_steps != 2) return fail(MLMCB_EINVAL, "packed levels have 1 or 2 fine steps, not %d", n_steps);
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    mlmcb_params p;
    memset(&p, 0, sizeof p);
    p.X0 = 1; p.K = 1; p.T = 1; p.sig = 1;
    LevelConsts C;
    if (int rc = make_consts(&p, MLMCB_GBM, MLMCB_EURO, 0, 2, n_paths, seed, path_offset, &C)) return rc;
    C.ctr3_tag = (uint32_t)l << 16;
    if (n_paths == 0) return 0;
    const int grid = (int)((n_paths + kThreads - 1) / kThreads);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (sampler == MLMCB_SAMPLER_F32BM) packed_probe_kernel<SAMP_F32BM><<<grid, kThreads, 0, st>>>(C, n_steps, d_out);
    else if (sampler == MLMCB_SAMPLER_F64BM) packed_probe_kernel<SAMP_F64BM><<<grid, kThreads, 0, st>>>(C, n_steps, d_out);
    else return fail(MLMCB_EINVAL, "unknown sampler %d", sampler);
    CU(cudaGetLastError());
    return 0;
}

int mlmcb_jump_probe(int sampler, uint64_t seed, int l, double lam, uint64_t path_offset, uint64_t n_paths,
                     uint64_t n_records, int device, void* cuda_stream, double* d_out) {
    if (!d_out) return fail(MLMCB_EINVAL, "d_out is NULL");
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    mlmcb_params p;
    memset(&p, 0, sizeof p);
    p.X0 = 1; p.K = 1; p.T = 1; p.sig = 1; p.lam = lam;
    LevelConsts C;
    if (int rc = make_consts(&p, MLMCB_MERTON, MLMCB_EURO, 0, 2, n_paths, seed, path_offset, &C)) return rc;
    C.ctr3_tag = (uint32_t)l << 16;
    if (n_paths == 0 || n_records == 0) return 0;
    const int grid = (int)((n_paths + kThreads - 1) / kThreads);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (sampler == MLMCB_SAMPLER_F32BM) jump_probe_kernel<SAMP_F32BM><<<grid, kThreads, 0, st>>>(C, n_records, d_out);
    else if (sampler == MLMCB_SAMPLER_F64BM) jump_probe_kernel<SAMP_F64BM><<<grid, kThreads, 0, st>>>(C, n_records, d_out);
    else return fail(MLMCB_EINVAL, "unknown sampler %d", sampler);
    CU(cudaGetLastError());
    return 0;
}

int mlmcb_launch_geometry(int model, int payoff, int M, int flags, int device, int* threads_per_block,
                          int* max_blocks) {
    if (int rc = check_flags(flags)) return rc;
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    int grid = 0, mb = 0;
    if (flags & MLMCB_ANTITHETIC) {
        if (int rc = check_anti(model, payoff, M)) return rc;
    }
    if (payoff == MLMCB_AMERICAN) {
        if (int rc = geometry(amer_level_kernel<SAMP_F32BM>, device, ~0ull >> 16, &grid, &mb)) return rc;
    } else {
        LevelKernel k = (flags & MLMCB_ANTITHETIC) ? pick_anti_kernel(payoff, M, flags) : pick_kernel(model, payoff, M, flags, ~0ull);
        if (int rc = geometry(k, device, ~0ull >> 16, &grid, &mb)) return rc;
    }
    if (threads_per_block) *threads_per_block = kThreads;
    if (max_blocks) *max_blocks = mb;
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Pipe micro-benchmarks: 8 independent dependent-chains per thread of one instruction kind.
// They give the roofline denominators MEASURED_PEAKS.json lacks (FP64 FMA issue rate) and the
// per-pipe rates the kernel's instruction mix is budgeted against.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kProbeIters = 2048, kProbeChains = 8, kProbeUnroll = 4;

template <int WHICH>
__global__ void __launch_bounds__(256) pipe_probe_kernel(float seed, unsigned* sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    double d[kProbeChains];
    float f[kProbeChains];
    unsigned u[kProbeChains];
    unsigned long long w[kProbeChains];
    unsigned f2u[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) {
        f2u[c] = tid + 3 * c;
        d[c] = 1.0 + 1e-9 * (tid + c); f[c] = 1.0f + seed * (tid + c); u[c] = tid * 2654435761u + c; w[c] = u[c];
    }
    const double da = 1.0000000001, db = 1e-12;
    const float fa = 1.0000001f, fb = seed;
    for (int it = 0; it < kProbeIters; ++it) {
#pragma unroll
        for (int r = 0; r < kProbeUnroll; ++r) {
#pragma unroll
            for (int c = 0; c < kProbeChains; ++c) {
                if (WHICH == 0) d[c] = fma(d[c], da, db);
                else if (WHICH == 1) d[c] = d[c] + db;
                else if (WHICH == 2) f[c] = fmaf(f[c], fa, fb);
                else if (WHICH == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]), "r"(0xD2511F53u));
                else if (WHICH == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                else if (WHICH == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
                else if (WHICH == 6) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(__float_as_uint(f[c])));
                else if (WHICH == 7) asm volatile("{ .reg .f64 t; cvt.f64.f32 t, %0; cvt.rn.f32.f64 %0, t; }" : "+f"(f[c]));
                else if (WHICH == 8) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
                // two-instruction mixes (independent chains): what a pair costs when both kinds are in flight
                else if (WHICH == 9) {       // DFMA + LOP3
                    d[c] = fma(d[c], da, db);
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                } else if (WHICH == 10) {    // DFMA + IMAD.WIDE
                    d[c] = fma(d[c], da, db);
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]), "r"(0xD2511F53u));
                } else if (WHICH == 11) {    // IMAD.WIDE + LOP3
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]), "r"(0xD2511F53u));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                } else if (WHICH == 12) {    // FFMA + LOP3
                    f[c] = fmaf(f[c], fa, fb);
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                } else if (WHICH == 13) {    // DFMA + FFMA
                    d[c] = fma(d[c], da, db);
                    f[c] = fmaf(f[c], fa, fb);
                } else if (WHICH == 14) {    // IMAD (32-bit multiply-add, not wide) + DFMA
                    d[c] = fma(d[c], da, db);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(0x9E3779B1u), "r"(tid));
                } else if (WHICH == 15) {    // IMAD.HI (high half of a 32x32 multiply)
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(0xD2511F53u), "r"(tid));
                } else if (WHICH == 16) {    // DFMA + IMAD.HI
                    d[c] = fma(d[c], da, db);
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(0xD2511F53u), "r"(tid));
                } else if (WHICH == 17) {    // IMAD.HI + LOP3
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(0xD2511F53u), "r"(tid));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(f2u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                } else if (WHICH == 18) {    // IMAD.HI + IMAD (both halves as two instructions)
                    asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(f2u[c]) : "r"(u[c]), "r"(0xD2511F53u), "r"(tid));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(0xD2511F53u), "r"(f2u[c]));
                }
            }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) acc ^= (unsigned)d[c] ^ __float_as_uint(f[c]) ^ u[c] ^ (unsigned)w[c] ^ f2u[c];
    if (acc == 0x12345u) sink[0] = acc;
}

// The headline loop's instruction mix per fine step (16 FP64, 6 IMAD.WIDE, ~15 INT32) issued two ways, on independent
// chains so that only the pipes and the issue port limit it:
//   MODE 0  every warp issues the whole mix, finely interleaved (what a perfect single-stream schedule could reach);
//   MODE 1  warp-specialised: of every three warps two issue the FP64 part (+ 8 INT32) and one issues the Philox part
//           (12 IMAD.WIDE + 14 INT32 = the draws of both) - the scheduler interleaves them at issue time.
// One block per SM; MODE 0: 8 warps (two per SMSP), MODE 1: 12 warps (two FP64 + one Philox warp per SMSP).
template <int MODE>
__global__ void __launch_bounds__(384) mix_probe_kernel(float seed, unsigned* sink, int iters) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, warp = threadIdx.x >> 5;
    double d[16];
    unsigned u[16];
    unsigned long long w[12];
#pragma unroll
    for (int c = 0; c < 16; ++c) { d[c] = 1.0 + 1e-9 * (tid + c); u[c] = tid * 2654435761u + c; }
#pragma unroll
    for (int c = 0; c < 12; ++c) w[c] = u[c];
    const double da = 1.0000000001, db = 1e-12;
    const bool philox_warp = MODE == 1 && warp >= 8;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                d[c] = fma(d[c], da, db);
                if (c < 15) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
                if (c < 6) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]), "r"(0xD2511F53u));
            }
        } else if (!philox_warp) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                d[c] = fma(d[c], da, db);
                if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
            }
        } else {
#pragma unroll
            for (int c = 0; c < 14; ++c) {
                if (c < 12) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]), "r"(0xD2511F53u));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(tid), "r"(0x9E3779B9u + it));
            }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc ^= (unsigned)d[c] ^ u[c];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc ^= (unsigned)w[c];
    if (acc == 0x12345u) sink[0] = acc;
}

typedef void (*ProbeKernel)(float, unsigned*);
ProbeKernel probe_of(int which) {
    switch (which) {
    case 0: return pipe_probe_kernel<0>;
    case 1: return pipe_probe_kernel<1>;
    case 2: return pipe_probe_kernel<2>;
    case 3: return pipe_probe_kernel<3>;
    case 4: return pipe_probe_kernel<4>;
    case 5: return pipe_probe_kernel<5>;
    case 6: return pipe_probe_kernel<6>;
    case 7: return pipe_probe_kernel<7>;
    case 8: return pipe_probe_kernel<8>;
    case 9: return pipe_probe_kernel<9>;
    case 10: return pipe_probe_kernel<10>;
    case 11: return pipe_probe_kernel<11>;
    case 12: return pipe_probe_kernel<12>;
    case 13: return pipe_probe_kernel<13>;
    case 14: return pipe_probe_kernel<14>;
    case 15: return pipe_probe_kernel<15>;
    case 16: return pipe_probe_kernel<16>;
    case 17: return pipe_probe_kernel<17>;
    case 18: return pipe_probe_kernel<18>;
    default: return nullptr;
    }
}

}  // namespace

// milliseconds of `iters` passes of the mix probe on every SM (one block each)
extern "C" int mlmcb_mix_probe(int mode, int device, int iters, float* ms_out) {
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    DeviceInfo* d;
    if (int rc = device_info(device, &d)) return rc;
    unsigned* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0;
    cudaError_t e = cudaMalloc((void**)&sink, 4);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) {
        if (rep) e = cudaEventRecord(e0);
        if (mode == 0) mix_probe_kernel<0><<<d->sms, 256>>>(1e-7f, sink, iters);
        else mix_probe_kernel<1><<<d->sms, 384>>>(1e-7f, sink, iters);
    }
    if (e == cudaSuccess) e = cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (e != cudaSuccess) return fail(MLMCB_ECUDA, "mix probe failed: %s", cudaGetErrorString(e));
    if (ms_out) *ms_out = ms;
    return 0;
}

extern "C" int mlmcb_pipe_probe(int which, int device, double* inst_per_s, float* ms_out) {
    ProbeKernel k = probe_of(which);
    if (!k) return fail(MLMCB_EINVAL, "unknown probe %d", which);
    DeviceGuard g(device);
    if (!g.ok) return fail(MLMCB_ECUDA, "cannot select CUDA device %d (no CPU fallback exists)", device);
    DeviceInfo* d;
    if (int rc = device_info(device, &d)) return rc;
    unsigned* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0;
    const int grid = d->sms * 8, threads = 256;
    cudaError_t e = cudaMalloc((void**)&sink, 4);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        k<<<grid, threads>>>(1e-7f, sink);   // warm-up
        e = cudaEventRecord(e0);
    }
    if (e == cudaSuccess) {
        k<<<grid, threads>>>(1e-7f, sink);
        e = cudaEventRecord(e1);
    }
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (e != cudaSuccess) return fail(MLMCB_ECUDA, "pipe probe failed: %s", cudaGetErrorString(e));
    // F2F probe issues two conversions per chain step, the mixes two instructions
    const double per_step = (which == 7 || (which >= 9 && which != 15)) ? 2.0 : 1.0;
    const double warp_insts = (double)grid * (threads / 32) * kProbeIters * kProbeUnroll * kProbeChains * per_step;
    if (inst_per_s) *inst_per_s = warp_insts / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}
