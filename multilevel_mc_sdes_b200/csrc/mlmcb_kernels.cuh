// mlmcb_kernels.cuh - device code of the B200 per-level coupled sampler (sm_100a).
//
// One thread simulates one coupled fine/coarse Euler-Maruyama path at a time, entirely in
// registers: Philox4x32-10 counter-based draws -> normals -> fine step every draw, coarse
// step every M draws with the summed Brownian increment -> path functional (final / trapezoid
// average / running minimum) -> payoff -> the seven looper sums, reduced warp-shuffle ->
// shared memory -> one row per block -> last-arriving block folds the rows in a fixed order.
// No path is ever written to HBM.
//
// What each piece restates (line numbers = /root/reference/MLMC_RSCAM.py):
//   gbm_*      diffusion_path :320-349, diffusion_path_avg :351-386, diffusion_path_min :388-426
//              with GBM_Option.sde :1363
//   JD<...>    JD_path :428-492, JD_path_avg :494-571, JD_path_min :573-651 (jump-ADAPTED EM)
//              with Merton_Option.sde :1314-1315
//   payoff_*   EA_payoff :62-86, Lookback_payoff :89-112, Digital_payoff :114-136
//   seven sums Option.looper :946-955
//
// Two arithmetic policies share the control flow:
//   AR_FAST   production form  X <- fma(X, fma(b,z,a), X)   (3 + 2/M FP64 issue slots / step)
//   AR_EXACT  the reference's left-to-right expression order with explicitly rounded
//             __dmul_rn/__dadd_rn (never contracted) - replay mode, bit-exact vs numpy for GBM.
#pragma once
#ifndef MLMCB_PACKED_F32
#define MLMCB_PACKED_F32 0          // FFMA2 costs two dispatch cycles on B200: no gain (profiles/r01_pipe_mix.txt)
#endif
#ifndef MLMCB_UNROLL
#define MLMCB_UNROLL 2              // two Philox blocks in flight per thread (measured best with 8 blocks/SM)
#endif
#ifndef MLMCB_MINBLOCKS
#define MLMCB_MINBLOCKS 8           // GBM/f32bm kernels: cap at 64 registers -> 8 blocks of 128 threads per SM
#endif
#ifndef MLMCB_MINBLOCKS_MJD
#define MLMCB_MINBLOCKS_MJD 6
#endif
#ifndef MLMCB_MJD_FLAT_BELOW
#define MLMCB_MJD_FLAT_BELOW 128    // Merton: levels with fewer fine steps run the event-driven kernel
#endif
#ifndef MLMCB_MINBLOCKS_FLAT
#define MLMCB_MINBLOCKS_FLAT 5      // event-driven Merton kernel: 96 registers
#endif
#ifndef MLMCB_JUMP_RECS
#define MLMCB_JUMP_RECS 4           // Merton: jump records produced up front per path (shared memory, 192 B per thread)
#endif
#ifndef MLMCB_UNROLL_MJD
#define MLMCB_UNROLL_MJD 2
#endif
#ifndef MLMCB_MINBLOCKS_F64
#define MLMCB_MINBLOCKS_F64 2       // fp64-sampler GBM kernels: four octets in flight (236 registers) at two blocks per SM beat
                                    // every higher-occupancy build (profiles/r02_f64bm_variants.txt: 66.4 vs 75.0 cycles per step)
#endif
#ifndef MLMCB_MINBLOCKS_F64_SHORT
#define MLMCB_MINBLOCKS_F64_SHORT 5 // fp64-sampler GBM packed / four-step kernels and the sweep kernel (where those levels hold nearly
                                    // all samples): one octet in flight, <= 96 registers.  (Measured: as a standalone kernel for 8..32
                                    // steps this build is 4-18 % SLOWER than the four-octet one, so level calls never use it.)
#endif
#ifndef MLMCB_UNROLL_F64
#define MLMCB_UNROLL_F64 4          // fp64 sampler: octets (three Philox calls, eight step factors) in flight per thread
#endif
#ifndef MLMCB_MINBLOCKS_F64_MJD
#define MLMCB_MINBLOCKS_F64_MJD 4   // Merton kernels with the fp64 sampler: <= 128 registers
#endif
#ifndef MLMCB_UNROLL_FW
#define MLMCB_UNROLL_FW 1           // Merton factor walk: jump-free octets in flight per thread (2 measured equal at 128 registers)
#endif
#ifndef MLMCB_FW_MIXED_FROM
#define MLMCB_FW_MIXED_FROM 32      // Merton factor walk: from this many fine steps mixed octets let the jump lanes walk alone (8, 16 steps: 2 %,
#endif                              // nearly every lane has a jump in every octet there and all lanes stay busy step by step)
#ifndef MLMCB_FW_SLOTS
#define MLMCB_FW_SLOTS 128          // Merton factor walk: record slots per warp
#endif
#ifndef MLMCB_PIPE_G
#define MLMCB_PIPE_G 2             // octets per stage of the software-pipelined fp64 GBM loop (the long-path kernel)
#endif
#ifndef MLMCB_PIPE_SHORT
#define MLMCB_PIPE_SHORT 1         // the same pipeline across paths in the four-step kernel (1-3 %)
#endif
#ifndef MLMCB_PIPE_FROM
#define MLMCB_PIPE_FROM 1024       // fine steps from which GBM levels take the long-path kernel (64 steps: 20 % slower, 256: equal, 1024+: 2-3 % faster)
#endif
#ifndef MLMCB_PHILOX_ROUNDS
#define MLMCB_PHILOX_ROUNDS 10      // experiments only; the shipped library is Philox4x32-10
#endif
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlmcb {

typedef unsigned long long ull;

enum { FN_FINAL = 0, FN_AVG = 1, FN_MIN = 2 };
enum { AR_FAST = 0, AR_EXACT = 1 };
enum { SAMP_F32BM = 0, SAMP_F64BM = 1 };
enum { SCH_EM = 0, SCH_MILSTEIN = 1 };
enum { PAY_EURO = 0, PAY_ASIAN = 1, PAY_LOOKBACK = 2, PAY_DIGITAL = 3, PAY_AMERICAN = 4 };

#ifndef MLMCB_THREADS
#define MLMCB_THREADS 128
#endif
constexpr int kThreads = MLMCB_THREADS;
constexpr int kUnrollMjd = MLMCB_UNROLL_MJD;
constexpr int kUnrollFw = MLMCB_UNROLL_FW;
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

// Everything a launch needs, computed once on the host (mlmcb_api.cu: make_consts) and passed
// by value as a __grid_constant__ kernel parameter (constant bank -> uniform operands).
struct LevelConsts {
    ull n_paths, path_offset, nsteps;
    int M, l, payoff, is_l0;
    uint32_t k0[10], k1[10];        // Philox round keys (key schedule hoisted to the host)
    uint32_t ctr3_tag;              // level << 16 ; ORed onto the high half of the path id
    // reference-order scalars (AR_EXACT) - each evaluated on the host exactly as numpy does
    double X0, K, T, mu, sig, dt, sqrt_dt, Mdt, Md;
    double disc, discK, look_cf, look_cc;
    // AR_FAST per-step coefficients:  t = a + b*z ;  X += X*t
    double a_f, b_f, a_c;
    double avg_wf, avg_wc;          // dt/T and M*dt/T
    // Merton
    double lam, inv_lam, jm, js;
    double inv_blk_dt;              // 1/(4*dt): jump time -> block index
    double inv_sqrt_dt;             // factor walk: rho = sqrt(d2) / sqrt(dt)
    double inv_T;                   // Merton average: Af/T as a multiplication (production arithmetic)
    uint32_t run_blocks;            // event-driven Merton walk: iterations of the advance phase between event phases
    double* raw;                    // optional [rows][n_paths] path-function outputs (Option.path), else NULL
    double r, boundary_c, amer_cneg;  // American put: rate, boundary constant, -2/(sig^2 dt); discounting inside the path, exercise boundary constant
    // Milstein scheme (ExampleUsages.ipynb cell 1: + 0.5*X*sig**2*(dW**2-dt))
    double sig2, half_sig2, a_fm, a_cm, c_mil;
    float poly_q4, poly_c4;         // leading sin/cos coefficients as uniform operands (no per-block MOV)
    // fp64 sampler constants, read straight from the constant bank by DFMA (no 64-bit immediates exist)
    double ln2_hi, ln2_lo;
    double lq[5], neg2ln2, ts[3], tc[2];   // table-driven fp64 Box-Muller (tools/gen_f64bm_tables.py)
    uint32_t one_hi;                // 0x3ff00000 as a run-time value: (x & mask) | one_hi then is ONE LOP3 (register operand)
    double ex[12];                  // 1/k!, k = 2..13 (exp_bounded)
    const double* amer_vals;        // HOST pointer (never read on the device): tabulated exercise boundary of this level, or null
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Round keys come precomputed; one round is two
// 32x32->64 multiplies (IMAD.WIDE.U32) and two three-input XORs (LOP3).
// ------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox_round(U4 c, uint32_t k0, uint32_t k1) {
    ull p0 = (ull)PHILOX_M0 * c.x, p1 = (ull)PHILOX_M1 * c.z;
    U4 r;
    r.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    r.y = (uint32_t)p1;
    r.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    r.w = (uint32_t)p0;
    return r;
}

__device__ __forceinline__ U4 philox10(const LevelConsts& C, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    U4 c = {c0, c1, c2, c3};
#pragma unroll
    for (int r = 0; r < MLMCB_PHILOX_ROUNDS; ++r) c = philox_round(c, C.k0[r], C.k1[r]);
    return c;
}

// Counter layout (128 bit):  c0 = stream kind (grid steps / jump stream)
//                            c1 = draw-block index (32 bit: 4 steps per block, M**l <= 2^33)
//                            c2,c3 = global path id (48 bit) | level (8 bit) | 0
// key = 64-bit seed.  Distinct (seed, level, path id, kind, block) -> distinct counter, so
// repeated looper calls (advancing path_offset), levels, ranks and the Merton jump stream never
// reuse a draw.  The block index sits in c1 - a word the first round only XORs - so that, with
// everything else fixed along a path, both multiplies of round 1 and one multiply each of rounds
// 2 and 3 are loop-invariant and hoisted by the compiler: 16 instead of 20 IMAD.WIDE per block
// for the same Philox4x32-10 output.
enum { STREAM_GRID = 0, STREAM_JUMP = 1, STREAM_PACK = 2, STREAM_GAP = 3 };

__device__ __forceinline__ U4 philox_at(const LevelConsts& C, ull pid, uint32_t kind, uint32_t blk) {
    return philox10(C, kind, blk, (uint32_t)pid, ((uint32_t)(pid >> 32) & 0xFFFFu) | C.ctr3_tag);
}

// ------------------------------------------------------------------------------------------
// Normal samplers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {   // MUFU.LG2, no denormal fix-up (x >= 2^-33)
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// fp32 Box-Muller.  Radius: u = (w+0.5)*2^-32 in (0,1] from a full 32-bit word (tail reaches
// 6.7 sigma), r = sqrt(-2 ln u) with MUFU.LG2 + MUFU.SQRT.  Angle: theta = pi*h, h in [-1/2,1/2)
// from a signed 32-bit word, cos/sin by degree-8/9 minimax polynomials in h (abs error 2e-7, no
// range reduction, no MUFU), and one spare random bit flips the sign of r so (r cos, r sin)
// covers the whole circle.  21 instructions per pair, 2 of them on the XU pipe.
#define MLMCB_Q0 3.1415927410125732f
#define MLMCB_Q1 -5.167710304260254f
#define MLMCB_Q2 2.5500776767730713f
#define MLMCB_Q3 -0.5982921719551086f
#define MLMCB_Q4 0.07765940576791763f
#define MLMCB_C0 0.9999999403953552f
#define MLMCB_C1 -4.934792995452881f
#define MLMCB_C2 4.058412075042725f
#define MLMCB_C3 -1.3318802118301392f
#define MLMCB_C4 0.2196968048810959f
#define MLMCB_NEG2LN2 -1.3862943611198906f

__device__ __forceinline__ void bm_pair_f32(const float q4, const float c4, uint32_t ru, uint32_t ra, float& z0, float& z1) {
    // ru as a signed word: |x| gives the uniform (31 bits + the 0.5 offset), sign(x) the half-plane
    const float x = __int2float_rn((int)ru);
    const float u = fmaf(fabsf(x), 0x1p-31f, 0x1p-32f);              // (0,1]
    float rad = sqrt_approx(lg2_approx(u) * MLMCB_NEG2LN2);         // sqrt(-2 ln u)
    rad = __uint_as_float((__float_as_uint(rad) & 0x7fffffffu) | (__float_as_uint(x) & 0x80000000u));   // one LOP3
    float h = __int2float_rn((int)ra) * 0x1p-32f;                   // angle / pi in [-1/2,1/2)
    float T = h * h;
    float q = fmaf(T, q4, MLMCB_Q3), c = fmaf(T, c4, MLMCB_C3);
    q = fmaf(q, T, MLMCB_Q2); c = fmaf(c, T, MLMCB_C2);
    q = fmaf(q, T, MLMCB_Q1); c = fmaf(c, T, MLMCB_C1);
    q = fmaf(q, T, MLMCB_Q0); c = fmaf(c, T, MLMCB_C0);
    z0 = rad * c;                                                    // r cos(pi h)
    z1 = (rad * h) * q;                                              // r sin(pi h)
}

// The same for two pairs at once with Blackwell's packed fp32 pipe (FFMA2/FMUL2): the two
// independent pairs of one Philox block ride in the .x/.y halves, halving the fp32 issue slots.
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

__device__ __forceinline__ void bm_quad_f32(const LevelConsts& C, U4 w, float z[4]) {
#if MLMCB_PACKED_F32 == 1
    float2 u = __ffma2_rn(make_float2(__uint2float_rn(w.x), __uint2float_rn(w.z)), f2(0x1p-32f), f2(0x1p-33f));
    float2 s = __fmul2_rn(make_float2(lg2_approx(u.x), lg2_approx(u.y)), f2(MLMCB_NEG2LN2));
    float2 rad = make_float2(sqrt_approx(s.x), sqrt_approx(s.y));
    rad.x = __uint_as_float(__float_as_uint(rad.x) ^ (w.x << 31));
    rad.y = __uint_as_float(__float_as_uint(rad.y) ^ (w.z << 31));
    float2 h = __fmul2_rn(make_float2(__int2float_rn((int)w.y), __int2float_rn((int)w.w)), f2(0x1p-32f));
    float2 T = __fmul2_rn(h, h);
    float2 q = __ffma2_rn(T, f2(MLMCB_Q4), f2(MLMCB_Q3)), c = __ffma2_rn(T, f2(MLMCB_C4), f2(MLMCB_C3));
    q = __ffma2_rn(q, T, f2(MLMCB_Q2)); c = __ffma2_rn(c, T, f2(MLMCB_C2));
    q = __ffma2_rn(q, T, f2(MLMCB_Q1)); c = __ffma2_rn(c, T, f2(MLMCB_C1));
    q = __ffma2_rn(q, T, f2(MLMCB_Q0)); c = __ffma2_rn(c, T, f2(MLMCB_C0));
    float2 zc = __fmul2_rn(rad, c), zs = __fmul2_rn(__fmul2_rn(rad, h), q);
    z[0] = zc.x; z[1] = zs.x; z[2] = zc.y; z[3] = zs.y;
#else
    bm_pair_f32(C.poly_q4, C.poly_c4, w.x, w.y, z[0], z[1]);
    bm_pair_f32(C.poly_q4, C.poly_c4, w.z, w.w, z[2], z[3]);
#endif
}

// ---- the fp64 Box-Muller sampler (`f64bm`, the default) --------------------------------------------
// One pair of N(0,1) from THREE Philox words (a, b, c):
//   radius  u = ((a & 0xFFFFF)*2^32 + b + 1/2) * 2^-52 in (0,1): 52 random bits; r = sqrt(-2 ln u), tail to 8.57 sigma
//   angle   theta = pi*h, h = h_j + d in [-1/2,1/2): j = bits 30..23 of a (h_j = (j + 1/2)/256 - 1/2),
//           d = c*2^-40 - 2^-9 (32 bits), and bit 31 of a picks the half-plane (sign of r): 41 angle bits
//   (z0, z1) = (r cos theta, r sin theta)
// so three Philox4x32-10 blocks give four pairs = eight normals.  All arithmetic is fp64, about 1 ulp:
//   -2 ln u = k(-2 ln2) + B_i + G(x),  u = 2^k m,  x = fma(m, A_i, 2) = -2(m/c_i - 1), G(x) = x + x^2 Q(x), |x| <= 2^-8,
//             256-entry table (A_i, B_i), Q of degree 3; the interval that holds m = 1 is centred on it (c = 1, A = -2, B = 0), so
//             x is exact there and u -> 1 keeps full relative accuracy;
//   sqrt      reciprocal-root seed (MUFU.RSQ64H) and one third-order correction;
//   sin/cos   256-entry table (cos, sin)(pi h_j), the remainder pi*d by three-/two-term polynomials, one rotation.
// 28 FP64-pipe instructions per pair (the round-1 table-free form - fdlibm log, degree-8 sin/cos - took 53).  The two
// tables (8 KB) are copied to shared memory by every kernel that draws fp64 normals.  tools/gen_f64bm_tables.py
// generates tables and coefficients and checks the pair against 60-digit values (worst |z - exact| 1.3e-15 at
// |z| = 8.6); tests/test_gpu_philox.py recomputes the device stream from the Philox words in numpy float64.

// (x & MASK) | c in one LOP3: MASK is the instruction's immediate, c must come from a register
template <uint32_t MASK>
__device__ __forceinline__ uint32_t and_or(uint32_t x, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "n"(MASK), "r"(c));
    return r;
}
__device__ __forceinline__ double rcp64_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));        // MUFU.RCP64H, ~2^-22
    return y;
}
__device__ __forceinline__ double rsqrt64_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // MUFU.RSQ64H, ~2^-22
    return y;
}
struct F64TabSrc { double2 lg[256]; double2 tr[256]; };
static __device__ const F64TabSrc g_f64tab = {
#include "mlmcb_f64tab.inc"
};
// In shared memory the two 4 KB tables start on a 4 KB boundary of the shared WINDOW: each then
// sits on a multiple of its own size, and "base + masked byte offset" is "base | masked offset" - folded into the
// LOP3 that masks.  The boundary is found at run time inside a buffer 4 KB larger than the tables (the first user
// variable of a block sits 1 KB into the window on this architecture, so `__align__` alone does not give it).
struct F64Tab { double2 tr[256]; double2 lg[256]; };
__device__ __forceinline__ uint32_t f64tab_tr() {
    __shared__ __align__(16) unsigned char buf[sizeof(F64Tab) + 4096];
    return ((uint32_t)__cvta_generic_to_shared(buf) + 4095u) & ~4095u;
}
__device__ __forceinline__ uint32_t f64tab_lg() { return f64tab_tr() + (uint32_t)sizeof(F64Tab::tr); }
// volatile: the statement has no declared memory operand, and only `volatile` keeps the compiler from moving it
// above the barrier that ends the table copy (an unordered asm was hoisted there and read garbage)
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" : : "r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
// every thread of the block, before the first fp64 draw
template <int SAMP>
__device__ __forceinline__ void sampler_init() {
    if (SAMP == SAMP_F64BM) {
        const uint32_t tr = f64tab_tr(), lg = f64tab_lg();
        for (int i = threadIdx.x; i < 256; i += blockDim.x) sts_f64x2(tr + 16u * i, g_f64tab.tr[i]);
        for (int i = threadIdx.x; i < 256; i += blockDim.x) sts_f64x2(lg + 16u * i, g_f64tab.lg[i]);
        __syncthreads();
    }
}
// (0,1) uniform from 52 random bits (low 20 of `hi`, all of `lo`) plus a half-ulp offset
__device__ __forceinline__ double u52(const LevelConsts& C, uint32_t hi, uint32_t lo) {
    return __hiloint2double((int)and_or<0x000FFFFFu>(hi, C.one_hi), (int)lo) - (1.0 - 0x1p-53);
}
// -2 ln u for a normal double u in (0,1)
__device__ __forceinline__ double neg2_log_u(const LevelConsts& C, double u) {
    const int ix = __double2hiint(u);
    const int tmp = ix - 0x3fe6b800;
    const int k = tmp >> 20;
    int mh;                                                                              // ix - (k << 20) as ONE IMAD
    asm("mad.lo.s32 %0, %1, -1048576, %2;" : "=r"(mh) : "r"(k), "r"(ix));
    const double m = __hiloint2double(mh, __double2loint(u));                            // [0.70947, 1.41895)
    const double2 e = lds_f64x2(and_or<0xff0u>((uint32_t)tmp >> 8, f64tab_lg()));                    // lg[(tmp >> 12) & 255]
    const double x = fma(m, e.x, 2.0);
    const double x2 = x * x;
    double q = fma(x, C.lq[3], C.lq[2]);
    q = fma(q, x, C.lq[1]);
    q = fma(q, x, C.lq[0]);
    const double t = fma((double)k, C.neg2ln2, e.y);
    return t + fma(x2, q, x);
}
// sqrt(s), s > 0 normal: reciprocal-root seed y (2^-22), g = s y, e = 1 - g y (exact to its own rounding), and
// sqrt(s) = g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8) with a truncation of 5 e^3/16 < 2^-64: five instructions
// (two residual corrections took six), < 1 ulp.
__device__ __forceinline__ double sqrt_pos(double s) {
    const double y = rsqrt64_seed(s);
    const double g = s * y;
    const double e = fma(-g, y, 1.0);
    return fma(g * e, fma(0.375, e, 0.5), g);
}
// AFFINE = false: (z0, z1);  AFFINE = true: (a0 + b0*z0, a0 + b0*z1) - the Euler-Maruyama step factor, one multiply
// per pair instead of one FMA per normal
template <bool AFFINE>
__device__ __forceinline__ void bm_pair_f64(const LevelConsts& C, uint32_t a, uint32_t b, uint32_t c, double& z0, double& z1,
                                            double a0 = 0.0, double b0 = 1.0) {
    double rad = sqrt_pos(neg2_log_u(C, u52(C, a, b)));
    rad = __hiloint2double(__double2hiint(rad) | (int)(a & 0x80000000u), __double2loint(rad));
    const double2 cs = lds_f64x2(and_or<0xff0u>(a >> 19, f64tab_tr()));        // tr[(a >> 23) & 255]
    const double dl = fma(__uint2double_rn(c), 0x1p-40, -0x1p-9);              // d in [-2^-9, 2^-9), exact
    const double D = dl * dl;
    const double sd = fma(fma(D, C.ts[2], C.ts[1]), D, C.ts[0]) * dl;          // sin(pi d)
    const double cm = fma(D, C.tc[1], C.tc[0]) * D;                             // cos(pi d) - 1
    const double cc = fma(-cs.y, sd, fma(cs.x, cm, cs.x)), ss = fma(cs.x, sd, fma(cs.y, cm, cs.y));
    if (AFFINE) {
        const double br = b0 * rad;
        z0 = fma(br, cc, a0);
        z1 = fma(br, ss, a0);
    } else {
        z0 = rad * cc;
        z1 = rad * ss;
    }
}

// max(x, 0) by masking with the sign: one shift and two logic operations.  (fp64 fmax - which is also what
// nvcc turns `x > 0 ? x : 0` into - expands to DSETP.MAX, two selects, a NaN-quieting LOP3 and moves.)
__device__ __forceinline__ double pos_part(double x) {
    const int hi = __double2hiint(x), keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}

// exp(x) for the jump size (:462), x clamped to [-700, 700] so the result is a normal number and the
// scaling is one integer add: n = rint(x/ln2) by the 1.5*2^52 shift, r = x - n*ln2 (two-part ln2, |r| <= 0.347),
// Taylor to r^13/13! (truncation 4e-18), 2^n into the exponent field.  <= 1 ulp; coefficients from the
// constant bank (libdevice's exp spends as many uniform moves on its constants as FMAs on the polynomial).
__device__ __forceinline__ double exp_bounded(const LevelConsts& C, double x) {
    x = x < -700.0 ? -700.0 : (x > 700.0 ? 700.0 : x);
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -C.ln2_hi, x);
    r = fma(nf, -C.ln2_lo, r);
    double p = fma(r, C.ex[11], C.ex[10]);
#pragma unroll
    for (int i = 9; i >= 0; --i) p = fma(p, r, C.ex[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

template <typename R> __device__ __forceinline__ R fma_(R a, R b, R c);
template <> __device__ __forceinline__ double fma_<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <typename R> __device__ __forceinline__ R min_(R a, R b);
template <> __device__ __forceinline__ double min_<double>(double a, double b) { return b < a ? b : a; }   // DSETP + 2 SEL; no NaN on this path
template <> __device__ __forceinline__ float min_<float>(float a, float b) { return fminf(a, b); }

// Four grid-step normals for (path, block b) = steps 4b+1 .. 4b+4.
// fp32 sampler: draw block b is Philox block b (four words, four normals).  fp64 sampler: a pair takes three
// words, so steps 8o+1 .. 8o+8 (an "octet") take Philox blocks 3o, 3o+1, 3o+2:
//   pair 0 (x,y,z of 3o) | pair 1 (w of 3o; x,y of 3o+1) | pair 2 (z,w of 3o+1; x of 3o+2) | pair 3 (y,z,w of 3o+2)
// load4(b) serves half an octet (o = b/2: two Philox calls), load8(o) a whole one (three calls).
template <typename R, int SAMP>
struct PhiloxSource {
    static constexpr int kUnroll = SAMP == SAMP_F32BM ? MLMCB_UNROLL : MLMCB_UNROLL_F64;   // draw blocks in flight per thread
    static constexpr bool kOctets = SAMP == SAMP_F64BM;
    const LevelConsts& C;
    ull pid;
    __device__ __forceinline__ void load4(uint32_t b, R z[4]) const { load4_stream(STREAM_GRID, b, z); }
    __device__ __forceinline__ void load4_stream(uint32_t kind, uint32_t b, R z[4]) const {
        if (SAMP == SAMP_F32BM) {
            float f[4];
            bm_quad_f32(C, philox_at(C, pid, kind, b), f);
            z[0] = (R)f[0]; z[1] = (R)f[1]; z[2] = (R)f[2]; z[3] = (R)f[3];
        } else {
            const uint32_t o3 = 3u * (b >> 1) + (b & 1u);
            const U4 r = philox_at(C, pid, kind, o3), s = philox_at(C, pid, kind, o3 + 1u);
            double d0, d1, d2, d3;
            if (b & 1u) {
                bm_pair_f64<false>(C, r.z, r.w, s.x, d0, d1);
                bm_pair_f64<false>(C, s.y, s.z, s.w, d2, d3);
            } else {
                bm_pair_f64<false>(C, r.x, r.y, r.z, d0, d1);
                bm_pair_f64<false>(C, r.w, s.x, s.y, d2, d3);
            }
            z[0] = (R)d0; z[1] = (R)d1; z[2] = (R)d2; z[3] = (R)d3;
        }
    }
    // the octet o as Euler-Maruyama step factors t_i = a0 + b0*z_i (fp64 sampler, fp64 paths)
    __device__ __forceinline__ void load8_affine(uint32_t o, double t[8], double a0, double b0) const {
        U4 w[3];
        words8(o, w);
        affine8(w, t, a0, b0);
    }
    // the first half of octet 0 (a four-step path) in the same two halves: two Philox blocks, two pairs -> four normals
    __device__ __forceinline__ void words4(U4 w[2], uint32_t kind = STREAM_GRID) const {
        w[0] = philox_at(C, pid, kind, 0u);
        w[1] = philox_at(C, pid, kind, 1u);
    }
    __device__ __forceinline__ void normals4(const U4 w[2], R z[4]) const {
        double d0, d1, d2, d3;
        bm_pair_f64<false>(C, w[0].x, w[0].y, w[0].z, d0, d1);
        bm_pair_f64<false>(C, w[0].w, w[1].x, w[1].y, d2, d3);
        z[0] = (R)d0; z[1] = (R)d1; z[2] = (R)d2; z[3] = (R)d3;
    }
    // the same in two halves - the integer half (three Philox blocks) and the FP64 half (four pairs) - so that a
    // loop can draw the words of its NEXT octets next to the floating-point work of the current ones
    __device__ __forceinline__ void words8(uint32_t o, U4 w[3]) const {
        w[0] = philox_at(C, pid, STREAM_GRID, 3u * o);
        w[1] = philox_at(C, pid, STREAM_GRID, 3u * o + 1u);
        w[2] = philox_at(C, pid, STREAM_GRID, 3u * o + 2u);
    }
    __device__ __forceinline__ void affine8(const U4 w[3], double t[8], double a0, double b0) const {
        bm_pair_f64<true>(C, w[0].x, w[0].y, w[0].z, t[0], t[1], a0, b0);
        bm_pair_f64<true>(C, w[0].w, w[1].x, w[1].y, t[2], t[3], a0, b0);
        bm_pair_f64<true>(C, w[1].z, w[1].w, w[2].x, t[4], t[5], a0, b0);
        bm_pair_f64<true>(C, w[2].y, w[2].z, w[2].w, t[6], t[7], a0, b0);
    }
};

// Normals of a path that shares a Philox block with its neighbours (coarsest levels): z[base..].
template <typename R>
struct PackedSource {
    static constexpr int kUnroll = 1;
    static constexpr bool kOctets = false;
    R z[4];
    int base;
    __device__ __forceinline__ void load4(uint32_t, R out[4]) const {   // two steps per path: base is 0 or 2
        out[0] = base == 0 ? z[0] : z[2];
        out[1] = base == 0 ? z[1] : z[3];
        out[2] = out[3] = (R)0;
    }
};

// Replayed normals, step-major Z[step*n + path] (MLMC_RSCAM.py:343 call order).
struct ReplaySource {
    static constexpr int kUnroll = 1;
    static constexpr bool kOctets = false;
    const double* Z;
    ull n, i, nsteps;
    __device__ __forceinline__ void load4(uint32_t b, double z[4]) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ull s = 4 * (ull)b + k;
            z[k] = s < nsteps ? Z[s * n + i] : 0.0;
        }
    }
};

// ------------------------------------------------------------------------------------------
// Payoffs and the seven sums.
// ------------------------------------------------------------------------------------------
// vf/vc: terminal value or average; gf/gc: running minima (lookback only).
__device__ __forceinline__ double payoff_fine(const LevelConsts& C, double v, double g) {
    switch (C.payoff) {
    case PAY_EURO:
    case PAY_ASIAN: return __dmul_rn(C.disc, pos_part(__dsub_rn(v, C.K)));                  // :79-80
    case PAY_LOOKBACK: return __dmul_rn(C.disc, __dsub_rn(v, __dmul_rn(g, C.look_cf)));      // :105-106
    default: return v > C.K ? C.discK : 0.0;                                                 // :131
    }
}
__device__ __forceinline__ double payoff_coarse(const LevelConsts& C, double v, double g) {
    switch (C.payoff) {
    case PAY_EURO:
    case PAY_ASIAN: return __dmul_rn(C.disc, pos_part(__dsub_rn(v, C.K)));                  // :84-85
    case PAY_LOOKBACK: return __dmul_rn(C.disc, __dsub_rn(v, __dmul_rn(g, C.look_cc)));      // :110-111
    default: return v > C.K ? C.discK : 0.0;                                                 // :135
    }
}

struct Sums7 {
    double s[7];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 7; ++k) s[k] = 0.0;
    }
    // Option.looper :946-955 ; l==0 row :949
    __device__ __forceinline__ void add(double Pf, double Pc, int is_l0) {
        if (is_l0) {
            s[0] += Pf; s[1] = fma(Pf, Pf, s[1]); s[2] += Pf; s[3] = fma(Pf, Pf, s[3]);
        } else {
            double d = Pf - Pc;
            s[0] += d; s[1] = fma(d, d, s[1]);
            s[2] += Pf; s[3] = fma(Pf, Pf, s[3]);
            s[4] += Pc; s[5] = fma(Pc, Pc, s[5]);
            s[6] = fma(Pc, Pf, s[6]);
        }
    }
};

// Payoff + sums + optional per-path outputs for the production kernels.  The path functional fixes the
// payoff formula (average -> EA payoff, minimum -> lookback, final value -> Euro or Digital), so only
// the last distinction is a run-time test.  (Hoisting the "any per-path output requested" test out of the
// path loop was measured: it costs the 4096-step loop 1 % through register allocation.)
template <int FN, bool MAYBE_L0 = true>
__device__ __forceinline__ void emit_t(const LevelConsts& C, ull i, double vf, double vc, double gf, double gc,
                                       Sums7& acc, double* Pf_out, double* Pc_out) {
    double Pf, Pc;
    if (FN == FN_MIN) {                                                  // :105-106, :110-111
        Pf = __dmul_rn(C.disc, __dsub_rn(vf, __dmul_rn(gf, C.look_cf)));
        Pc = __dmul_rn(C.disc, __dsub_rn(vc, __dmul_rn(gc, C.look_cc)));
    } else if (FN == FN_AVG || C.payoff != PAY_DIGITAL) {                // :79-80, :84-85
        Pf = __dmul_rn(C.disc, pos_part(__dsub_rn(vf, C.K)));
        Pc = __dmul_rn(C.disc, pos_part(__dsub_rn(vc, C.K)));
    } else {                                                             // :131, :135
        Pf = vf > C.K ? C.discK : 0.0;
        Pc = vc > C.K ? C.discK : 0.0;
    }
    if (MAYBE_L0 && C.is_l0) Pc = vc;                                    // l==0: raw coarse output :82
    acc.add(Pf, Pc, MAYBE_L0 && C.is_l0);
    if (Pf_out || Pc_out || C.raw) {
        if (Pf_out) Pf_out[i] = Pf;
        if (Pc_out) Pc_out[i] = Pc;
        if (C.raw) {  // what the reference path function returns: (Xf,Xc) | (Af/T,Ac/T) | (Xf,Mf,Xc,Mc)
            C.raw[i] = vf; C.raw[C.n_paths + i] = gf; C.raw[2 * C.n_paths + i] = vc; C.raw[3 * C.n_paths + i] = gc;
        }
    }
}

// Block-level fold + deterministic cross-block finish.  `partials` has gridDim.x rows of 8
// doubles; `ticket` is zero at launch.  The last block to arrive sums the rows in index order.
// block_fold returns, in thread k < 7, the block total of sum k (other threads: 0).
__device__ __forceinline__ double block_fold(const Sums7& a, double* sm /* [kThreads/32][8] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double v = a.s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[w * 8 + k] = v;
    }
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 7)
        for (int j = 0; j < (int)(blockDim.x >> 5); ++j) r += sm[j * 8 + threadIdx.x];
    return r;
}

__device__ __forceinline__ void grid_finish(const Sums7& a, double* partials, unsigned* ticket, double* out7) {
    __shared__ double sm[(kThreads / 32) * 8];
    __shared__ unsigned last;
    const double mine = block_fold(a, sm);
    if (threadIdx.x < 7) partials[(size_t)blockIdx.x * 8 + threadIdx.x] = mine;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    Sums7 t;
    t.clear();
    for (unsigned r = threadIdx.x; r < gridDim.x; r += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 7; ++k) t.s[k] += __ldcg(&partials[(size_t)r * 8 + k]);
    }
    const double tot = block_fold(t, sm);
    if (threadIdx.x < 7) out7[threadIdx.x] = tot;
}

// Per-step multiplicative update X <- X + X*t(z).  Euler-Maruyama: t = a + b z.  Milstein adds
// 0.5 sig^2 (dW^2 - dt) = c z^2 - c with c = 0.5 sig^2 dt, i.e. t = (a - c) + b z + c z^2; the
// coarse step sees dWc = sqrt(dt)*sum z over M dt:  t_c = (a_c - M c) + b s + c s^2.
template <typename R, int SCH>
struct StepCoef {
    R a, b, c, ac;
    __device__ __forceinline__ explicit StepCoef(const LevelConsts& C)
        : a((R)(SCH ? C.a_fm : C.a_f)), b((R)C.b_f), c((R)C.c_mil), ac((R)(SCH ? C.a_cm : C.a_c)) {}
    __device__ __forceinline__ R t(R z) const { return SCH ? fma_<R>(fma_<R>(c, z, b), z, a) : fma_<R>(b, z, a); }
    __device__ __forceinline__ R tc(R s) const { return SCH ? fma_<R>(fma_<R>(c, s, b), s, ac) : fma_<R>(b, s, ac); }
};

// ------------------------------------------------------------------------------------------
// GBM coupled path, production arithmetic.  Source supplies normals four steps at a time.
// MT = 2 or 4: coarse refinement factor known at compile time (coarse steps fall on fixed
// positions of a 4-step block); MT = 0: any M, a counter decides.
// Returns the path outputs the reference path function returns: terminal values (FN_FINAL),
// averages (FN_AVG) or terminal values + minima (FN_MIN).
// ------------------------------------------------------------------------------------------
template <typename R, int FN, int MT, int SCH, typename Source, int UNROLL = Source::kUnroll, bool PIPE = false>
__device__ __forceinline__ void gbm_path_fast(const LevelConsts& C, const Source& src,
                                              double& vf, double& vc, double& gf, double& gc) {
    if constexpr (Source::kOctets && sizeof(R) == 8 && SCH == SCH_EM && (MT == 2 || MT == 4)) {
        // fp64 sampler, Euler-Maruyama, M = 2 or 4: the draws arrive eight steps at a time (three Philox blocks) and
        // already as step factors t_i = a + b z_i, so a fine step is ONE fma and - because the coarse drift is M times
        // the fine one - the coarse factor a_c + b*sum(z) is just the sum of its fine factors.
        if ((C.nsteps & 7) == 0) {
            double Xf = C.X0, Xc = C.X0;
            double Ff = FN == FN_MIN ? C.X0 : 0.0, Fc = Ff;
            const uint32_t noct = (uint32_t)(C.nsteps >> 3);
            // PIPE (the long-path build: its own kernel, so that the plain loop's register allocation does not move):
            // software pipeline - the Philox words of the next group of octets are drawn (integer pipes) in the same
            // loop body as the Box-Muller pairs and steps of the current group (FP64 pipe); ptxas otherwise runs all
            // twelve Philox blocks of an iteration first and the FP64 work after them, and IMAD.WIDE + LOP3 pairs cost
            // 3.2 cycles where DFMA + LOP3 cost 2.4 (bench.py pipe_probe)
            constexpr int G = MLMCB_PIPE_G;
            uint32_t o0 = 0;            // octets done by the pipelined loop; what is left takes the plain one
            if constexpr (PIPE) {
              if (noct >= 2 * G) {
                const uint32_t npipe = noct - noct % (2 * G);
                U4 wa[G][3], wb[G][3];
#pragma unroll
                for (int q = 0; q < G; ++q) src.words8(q, wa[q]);
                auto consume = [&](const U4 (&w)[G][3]) {
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        double t[8];
                        src.affine8(w[q], t, C.a_f, C.b_f);
#pragma unroll
                        for (int g = 0; g < 8; g += MT) {
                            const double tc = MT == 4 ? (t[g] + t[g + 1]) + (t[g + 2] + t[g + 3]) : t[g] + t[g + 1];
#pragma unroll
                            for (int k = 0; k < MT; ++k) {
                                Xf = fma(Xf, t[g + k], Xf);
                                if (FN == FN_AVG) Ff += Xf;
                                if (FN == FN_MIN) Ff = min_<double>(Ff, Xf);
                            }
                            Xc = fma(Xc, tc, Xc);
                            if (FN == FN_AVG) Fc += Xc;
                            if (FN == FN_MIN) Fc = min_<double>(Fc, Xc);
                        }
                    }
                };
                uint32_t o = 0;
#pragma unroll 1
                for (; o + 2 * G < npipe; o += 2 * G) {
#pragma unroll
                    for (int q = 0; q < G; ++q) src.words8(o + G + q, wb[q]);
                    consume(wa);
#pragma unroll
                    for (int q = 0; q < G; ++q) src.words8(o + 2 * G + q, wa[q]);
                    consume(wb);
                }
#pragma unroll
                for (int q = 0; q < G; ++q) src.words8(o + G + q, wb[q]);                // the last pass draws nothing past the end
                consume(wa);
                consume(wb);
                o0 = npipe;
              }
            }
#pragma unroll UNROLL
            for (uint32_t o = o0; o < noct; ++o) {
                double t[8];
                src.load8_affine(o, t, C.a_f, C.b_f);
#pragma unroll
                for (int g = 0; g < 8; g += MT) {
                    const double tc = MT == 4 ? (t[g] + t[g + 1]) + (t[g + 2] + t[g + 3]) : t[g] + t[g + 1];
#pragma unroll
                    for (int k = 0; k < MT; ++k) {
                        Xf = fma(Xf, t[g + k], Xf);
                        if (FN == FN_AVG) Ff += Xf;
                        if (FN == FN_MIN) Ff = min_<double>(Ff, Xf);
                    }
                    Xc = fma(Xc, tc, Xc);
                    if (FN == FN_AVG) Fc += Xc;
                    if (FN == FN_MIN) Fc = min_<double>(Fc, Xc);
                }
            }
            if (FN == FN_AVG) {
                const double h0 = 0.5 * C.X0;
                vf = C.avg_wf * (Ff + fma(-0.5, Xf, h0));
                vc = C.avg_wc * (Fc + fma(-0.5, Xc, h0));
                gf = gc = 0.0;
            } else {
                vf = Xf; vc = Xc; gf = Ff; gc = Fc;
            }
            return;
        }
    }
    const StepCoef<R, SCH> co(C);
    R Xf = (R)C.X0, Xc = (R)C.X0, sz = (R)0;
    R Ff = FN == FN_MIN ? (R)C.X0 : (R)0, Fc = Ff;
    int cm = 0;
    const int M = C.M;

#define MLMCB_FINE(zz)                                                   \
    {                                                                    \
        R t_ = co.t(zz);                                    \
        Xf = fma_<R>(Xf, t_, Xf);                                        \
        sz += (zz);                                                      \
        if (FN == FN_AVG) Ff += Xf;                                      \
        if (FN == FN_MIN) Ff = min_<R>(Ff, Xf);                          \
    }
#define MLMCB_FINE_NOSUM(zz)                                             \
    {                                                                    \
        R t_ = co.t(zz);                                    \
        Xf = fma_<R>(Xf, t_, Xf);                                        \
        if (FN == FN_AVG) Ff += Xf;                                      \
        if (FN == FN_MIN) Ff = min_<R>(Ff, Xf);                          \
    }
#define MLMCB_COARSE()                                                   \
    {                                                                    \
        R t_ = co.tc(sz);                                      \
        Xc = fma_<R>(Xc, t_, Xc);                                        \
        sz = (R)0;                                                       \
        if (FN == FN_AVG) Fc += Xc;                                      \
        if (FN == FN_MIN) Fc = min_<R>(Fc, Xc);                          \
    }

    const uint32_t nblk = (uint32_t)(C.nsteps >> 2);
#pragma unroll UNROLL
    for (uint32_t b = 0; b < nblk; ++b) {
        R z[4];
        src.load4(b, z);
        if (MT == 4) {               // coarse increment as a tree, off the fine chain
            const R s4 = (z[0] + z[1]) + (z[2] + z[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) MLMCB_FINE_NOSUM(z[k]);
            sz = s4;
            MLMCB_COARSE();
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                MLMCB_FINE(z[k]);
                if (MT == 2) { if (k & 1) MLMCB_COARSE(); }
                else { if (++cm == M) { cm = 0; MLMCB_COARSE(); } }
            }
        }
    }
    const int rem = (int)(C.nsteps & 3);
    if (rem) {                       // l==0 (1 step), M=2,l=1 (2 steps), odd M
        R z[4];
        src.load4(nblk, z);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < rem) {
                MLMCB_FINE(z[k]);
                if (++cm == M) { cm = 0; MLMCB_COARSE(); }
            }
        }
    }
#undef MLMCB_FINE
#undef MLMCB_FINE_NOSUM
#undef MLMCB_COARSE

    if (FN == FN_AVG) {              // trapezoid: dt*(X0/2 + sum_{j<=N} X_j - X_N/2)/T   :369-386
        const double h0 = 0.5 * C.X0;
        vf = C.avg_wf * ((double)Ff + fma(-0.5, (double)Xf, h0));
        vc = C.avg_wc * ((double)Fc + fma(-0.5, (double)Xc, h0));
        gf = gc = 0.0;
    } else {
        vf = (double)Xf; vc = (double)Xc; gf = (double)Ff; gc = (double)Fc;
    }
}

// GBM coupled path, reference expression order (replay, bit-exact vs numpy).
// Milstein increment in the notebook's order: ... + 0.5*X*(sig**2)*(dW**2-dt)
__device__ __forceinline__ double milstein_term(const LevelConsts& C, double X, double h, double dW) {
    return __dmul_rn(__dmul_rn(__dmul_rn(0.5, X), C.sig2), __dsub_rn(__dmul_rn(dW, dW), h));
}

template <int FN, int SCH>
__device__ __forceinline__ void gbm_path_exact(const LevelConsts& C, const double* Z, ull n, ull i,
                                               double& vf, double& vc, double& gf, double& gc) {
    const double dt = C.dt, Mdt = C.Mdt, mu = C.mu, sig = C.sig, Md = C.Md;
    double Xf = C.X0, Xc = C.X0, dWc = 0.0;
    double Af = __dmul_rn(__dmul_rn(0.5, dt), Xf);                       // :369
    double Ac = __dmul_rn(__dmul_rn(__dmul_rn(0.5, Md), dt), Xc);        // :370
    double Mf = C.X0, Mc = C.X0;
    int cm = 0;
    for (ull j = 0; j < C.nsteps; ++j) {
        double dWf = __dmul_rn(Z[j * n + i], C.sqrt_dt);                 // :343
        dWc = __dadd_rn(dWc, dWf);                                       // :344
        {
            double inc = __dadd_rn(__dmul_rn(__dmul_rn(mu, Xf), dt), __dmul_rn(__dmul_rn(sig, Xf), dWf));
            if (SCH) inc = __dadd_rn(inc, milstein_term(C, Xf, dt, dWf));
            Xf = __dadd_rn(Xf, inc);                                     // :345
        }
        if (FN == FN_AVG) Af = __dadd_rn(Af, __dmul_rn(Xf, dt));         // :377
        if (FN == FN_MIN) Mf = fmin(Xf, Mf);                             // :420
        if (++cm == C.M) {
            cm = 0;
            {
                double inc = __dadd_rn(__dmul_rn(__dmul_rn(mu, Xc), Mdt), __dmul_rn(__dmul_rn(sig, Xc), dWc));
                if (SCH) inc = __dadd_rn(inc, milstein_term(C, Xc, Mdt, dWc));
                Xc = __dadd_rn(Xc, inc);                                 // :347
            }
            if (FN == FN_AVG) Ac = __dadd_rn(Ac, __dmul_rn(__dmul_rn(Xc, Md), dt));   // :380
            if (FN == FN_MIN) Mc = fmin(Xc, Mc);                         // :423
            dWc = 0.0;
        }
    }
    if (FN == FN_AVG) {
        Af = __dsub_rn(Af, __dmul_rn(__dmul_rn(0.5, Xf), dt));           // :383
        Ac = __dsub_rn(Ac, __dmul_rn(__dmul_rn(__dmul_rn(0.5, Xc), Md), dt));   // :384
        vf = __ddiv_rn(Af, C.T); vc = __ddiv_rn(Ac, C.T);                // :386
        gf = gc = 0.0;
    } else {
        vf = Xf; vc = Xc; gf = Mf; gc = Mc;
    }
}

// ------------------------------------------------------------------------------------------
// Antithetic GBM triple (fine, antithetic-fine, coarse): diffusion_anti_path :162-217 and
// diffusion_anti_path_avg :219-290.  The antithetic path swaps the two Brownian increments of
// every pair of fine steps; M must be even (:178).  FN is FN_FINAL or FN_AVG.
// ------------------------------------------------------------------------------------------
template <typename R, int FN, int MT, int SCH, typename Source>
__device__ __forceinline__ void gbm_anti_path_fast(const LevelConsts& C, const Source& src,
                                                   double& vf, double& va, double& vc) {
    const StepCoef<R, SCH> co(C);
    R Xf = (R)C.X0, Xa = (R)C.X0, Xc = (R)C.X0, sz = (R)0;
    R Ff = (R)0, Fa = (R)0, Fc = (R)0;
    int cm = 0;
    const int M = C.M;
    if (C.is_l0) {                                                       // :213-215 / :282-285
        R z[4];
        src.load4(0, z);
        R t_ = co.t(z[0]);
        Xf = fma_<R>(Xf, t_, Xf);
        vf = va = FN == FN_AVG ? C.avg_wf * (0.5 * C.X0 + 0.5 * (double)Xf) : (double)Xf;
        vc = C.X0;
        return;
    }
#define MLMCB_PAIR(zo, ze)                                               \
    {                                                                    \
        R to_ = co.t(zo), te_ = co.t(ze);                                \
        Xf = fma_<R>(Xf, to_, Xf); Xa = fma_<R>(Xa, te_, Xa);            \
        if (FN == FN_AVG) { Ff += Xf; Fa += Xa; }                        \
        Xf = fma_<R>(Xf, te_, Xf); Xa = fma_<R>(Xa, to_, Xa);            \
        if (FN == FN_AVG) { Ff += Xf; Fa += Xa; }                        \
        sz += (zo) + (ze);                                               \
    }
#define MLMCB_COARSE()                                                   \
    {                                                                    \
        R t_ = co.tc(sz);                                                \
        Xc = fma_<R>(Xc, t_, Xc);                                        \
        sz = (R)0;                                                       \
        if (FN == FN_AVG) Fc += Xc;                                      \
    }
    const uint32_t nblk = (uint32_t)(C.nsteps >> 2);
    for (uint32_t b = 0; b < nblk; ++b) {
        R z[4];
        src.load4(b, z);
        MLMCB_PAIR(z[0], z[1]);
        if (MT == 2) MLMCB_COARSE()
        else if (MT == 0) { cm += 2; if (cm == M) { cm = 0; MLMCB_COARSE(); } }
        MLMCB_PAIR(z[2], z[3]);
        if (MT == 2 || MT == 4) MLMCB_COARSE()
        else { cm += 2; if (cm == M) { cm = 0; MLMCB_COARSE(); } }
    }
    if (C.nsteps & 2) {                                                  // M=2,l=1 or M=6: one trailing pair
        R z[4];
        src.load4(nblk, z);
        MLMCB_PAIR(z[0], z[1]);
        cm += 2;
        if (cm == M) { cm = 0; MLMCB_COARSE(); }
    }
#undef MLMCB_PAIR
#undef MLMCB_COARSE
    if (FN == FN_AVG) {
        const double h0 = 0.5 * C.X0;
        vf = C.avg_wf * ((double)Ff + fma(-0.5, (double)Xf, h0));
        va = C.avg_wf * ((double)Fa + fma(-0.5, (double)Xa, h0));
        vc = C.avg_wc * ((double)Fc + fma(-0.5, (double)Xc, h0));
    } else {
        vf = (double)Xf; va = (double)Xa; vc = (double)Xc;
    }
}

// reference expression order (replay, bit-exact vs numpy)
template <int FN, int SCH>
__device__ __forceinline__ double anti_inc(const LevelConsts& C, double X, double h, double dW) {
    double inc = __dadd_rn(__dmul_rn(__dmul_rn(C.mu, X), h), __dmul_rn(__dmul_rn(C.sig, X), dW));
    if (SCH) inc = __dadd_rn(inc, milstein_term(C, X, h, dW));           // whatever `self.sde` is bound to (:200-210)
    return inc;
}
template <int FN, int SCH>
__device__ __forceinline__ void gbm_anti_path_exact(const LevelConsts& C, const double* Z, ull n, ull i,
                                                    double& vf, double& va, double& vc) {
    const double dt = C.dt, Mdt = C.Mdt, Md = C.Md;
#define MLMCB_INC(X, h, dW) anti_inc<FN, SCH>(C, (X), (h), (dW))
    double Xf = C.X0, Xa = C.X0, Xc = C.X0, dWc = 0.0;
    double Af = __dmul_rn(__dmul_rn(0.5, dt), Xf), Aa = __dmul_rn(__dmul_rn(0.5, dt), Xa);      // :247-248
    double Ac = __dmul_rn(__dmul_rn(__dmul_rn(0.5, Md), dt), Xc);                               // :249
    if (C.is_l0) {
        Xf = __dadd_rn(Xf, MLMCB_INC(Xf, dt, __dmul_rn(Z[i], C.sqrt_dt)));                       // :214
        if (FN == FN_AVG) { Af = __dadd_rn(Af, __dmul_rn(__dmul_rn(0.5, dt), Xf)); vf = __ddiv_rn(Af, C.T); }   // :284
        else vf = Xf;
        va = vf; vc = Xc;
        return;
    }
    int cm = 0;
    for (ull k = 0; k + 1 < C.nsteps; k += 2) {
        const double dWo = __dmul_rn(Z[k * n + i], C.sqrt_dt), dWe = __dmul_rn(Z[(k + 1) * n + i], C.sqrt_dt);   // :195-196
        Xf = __dadd_rn(Xf, MLMCB_INC(Xf, dt, dWo));                      // :200
        Xa = __dadd_rn(Xa, MLMCB_INC(Xa, dt, dWe));                      // :201
        if (FN == FN_AVG) { Af = __dadd_rn(Af, __dmul_rn(dt, Xf)); Aa = __dadd_rn(Aa, __dmul_rn(dt, Xa)); }    // :263-264
        Xf = __dadd_rn(Xf, MLMCB_INC(Xf, dt, dWe));                      // :205
        Xa = __dadd_rn(Xa, MLMCB_INC(Xa, dt, dWo));                      // :206
        if (FN == FN_AVG) { Af = __dadd_rn(Af, __dmul_rn(dt, Xf)); Aa = __dadd_rn(Aa, __dmul_rn(dt, Xa)); }    // :271-272
        dWc = __dadd_rn(dWc, __dadd_rn(dWo, dWe));                       // :208
        cm += 2;
        if (cm == C.M) {
            cm = 0;
            Xc = __dadd_rn(Xc, MLMCB_INC(Xc, Mdt, dWc));                 // :210
            if (FN == FN_AVG) Ac = __dadd_rn(Ac, __dmul_rn(__dmul_rn(dt, Md), Xc));      // :278
            dWc = 0.0;
        }
    }
#undef MLMCB_INC
    if (FN == FN_AVG) {
        Af = __dsub_rn(Af, __dmul_rn(__dmul_rn(0.5, Xf), dt));           // :287
        Aa = __dsub_rn(Aa, __dmul_rn(__dmul_rn(0.5, Xa), dt));           // :288
        Ac = __dsub_rn(Ac, __dmul_rn(__dmul_rn(0.5, Xc), Mdt));          // :289
        vf = __ddiv_rn(Af, C.T); va = __ddiv_rn(Aa, C.T); vc = __ddiv_rn(Ac, C.T);
    } else {
        vf = Xf; va = Xa; vc = Xc;
    }
}

// anti_EA_payoff :292-317
__device__ __forceinline__ void emit_anti(const LevelConsts& C, ull i, double vf, double va, double vc,
                                          Sums7& acc, double* Pf_out, double* Pc_out) {
    double Pf = payoff_fine(C, vf, 0.0), Pc;
    if (C.is_l0) {
        Pc = vc;                                                         // :311
    } else {
        Pf = __dmul_rn(0.5, __dadd_rn(Pf, payoff_fine(C, va, 0.0)));     // :317
        Pc = payoff_coarse(C, vc, 0.0);
    }
    acc.add(Pf, Pc, C.is_l0);
    if (Pf_out) Pf_out[i] = Pf;
    if (Pc_out) Pc_out[i] = Pc;
    if (C.raw) { C.raw[i] = vf; C.raw[C.n_paths + i] = va; C.raw[2 * C.n_paths + i] = vc; }   // anti_path :217/:290
}

// ------------------------------------------------------------------------------------------
// Merton jump-adapted coupled path state (JD_path* :428-651).  Times are always double.
// G = trapezoid integral (FN_AVG) or running minimum (FN_MIN).
// ------------------------------------------------------------------------------------------
template <typename R, int FN, int AR, int SCH = SCH_EM>
struct JD {
    R Sf, Sc, Gf, Gc, dWc;
    double dtc, tf, tc;

    __device__ __forceinline__ void init(const LevelConsts& C) {
        Sf = Sc = (R)C.X0;
        Gf = Gc = FN == FN_MIN ? (R)C.X0 : (R)0;
        dWc = (R)0;
        dtc = tf = tc = 0.0;
    }

    // diffusion part of Merton_Option.sde :1314  dx_ = c*X*dt + sig*X*dW
    __device__ __forceinline__ R dx(const LevelConsts& C, R X, double dt, R dW) const {
        if (AR == AR_EXACT) {
            double d = __dadd_rn(__dmul_rn(__dmul_rn(C.mu, (double)X), dt), __dmul_rn(__dmul_rn(C.sig, (double)X), (double)dW));
            if (SCH) d = __dadd_rn(d, milstein_term(C, (double)X, dt, (double)dW));
            return (R)d;
        }
        if (SCH) return X * fma_<R>(fma_<R>((R)C.half_sig2, dW, (R)C.sig), dW, (R)((C.mu - C.half_sig2) * dt));
        return X * fma_<R>((R)C.sig, dW, (R)(C.mu * dt));
    }
    __device__ __forceinline__ R half_area(R S, double dt) const {       // 0.5*dt*S :532
        if (AR == AR_EXACT) return (R)__dmul_rn(__dmul_rn(0.5, dt), (double)S);
        return (R)(0.5 * dt) * S;
    }
    __device__ __forceinline__ static R add(R a, R b) {
        if (AR == AR_EXACT) return (R)__dadd_rn((double)a, (double)b);
        return a + b;
    }
    __device__ __forceinline__ static R mul(R a, R b) {
        if (AR == AR_EXACT) return (R)__dmul_rn((double)a, (double)b);
        return a * b;
    }

    // one pass of the `while tau<tn` body :458-474 (the caller advances tau afterwards)
    __device__ __forceinline__ static R jump_size(const LevelConsts& C, R zq) {                          // :462
        if (AR == AR_EXACT) return (R)__dsub_rn(exp(__dadd_rn(C.jm, __dmul_rn(C.js, (double)zq))), 1.0);
        return (R)(exp_bounded(C, fma(C.js, (double)zq, C.jm)) - 1.0);
    }
    __device__ __forceinline__ void jump(const LevelConsts& C, double tau, R zw, R zq) {
        jump_dJ(C, tau, zw, jump_size(C, zq));
    }
    __device__ __forceinline__ void jump_dJ(const LevelConsts& C, double tau, R zw, R dJ) {
        double dt = tau - tf;                                            // :458
        R dWf = mul(zw, (R)(AR == AR_EXACT ? sqrt(dt) : (dt > 0.0 ? sqrt_pos(dt) : 0.0)));   // :460
        jump_pre(C, dt, dWf, dJ);
        tf = tau; tc = tau;                                              // :472-473
    }
    // the same with the step length and Brownian increment already formed (clocks untouched).
    // COARSE = false (l = 0 only): between jumps the coarse path of a one-step level does not move, and at
    // a jump it takes the very increments of the fine one (dtc = dt, dWc = dWf), so it IS the fine
    // state after the jump - copied instead of recomputed.
    template <bool COARSE = true>
    __device__ __forceinline__ void jump_pre(const LevelConsts& C, double dt, R dWf, R dJ) {
        if (COARSE) {
            dtc += dt;                                                   // :459
            dWc = add(dWc, dWf);                                         // :461
        }
        if (FN == FN_AVG) {                                              // :532-540
            Gf = add(Gf, half_area(Sf, dt));
            Sf = add(Sf, dx(C, Sf, dt, dWf));
            Gf = add(Gf, half_area(Sf, dt));
            Sf = add(Sf, mul(Sf, dJ));
            if (COARSE) {
                Gc = add(Gc, half_area(Sc, dtc));
                Sc = add(Sc, dx(C, Sc, dtc, dWc));
                Gc = add(Gc, half_area(Sc, dtc));
                Sc = add(Sc, mul(Sc, dJ));
            }
        } else {                                                         // :465-468, sde with dJ :1315
            R d = dx(C, Sf, dt, dWf);
            Sf = add(Sf, add(d, mul(add(d, Sf), dJ)));
            if (COARSE) {
                d = dx(C, Sc, dtc, dWc);
                Sc = add(Sc, add(d, mul(add(d, Sc), dJ)));
            }
            if (FN == FN_MIN) {                                          // :621-622
                Gf = min_<R>(Gf, Sf);
                if (COARSE) Gc = min_<R>(Gc, Sc);
            }
        }
        if (!COARSE) { Sc = Sf; Gc = Gf; }
        dWc = (R)0; dtc = 0.0;                                           // :470-471
    }

    // the remainder step up to the grid time tn :477-482 (+ :637).  `whole` (production policy
    // only): no jump fell inside this fine step, so dt is the grid step and its root is known.
    __device__ __forceinline__ void grid(const LevelConsts& C, double tn, R zw, bool whole = false) {
        const double dt = whole ? C.dt : tn - tf;                        // :477
        R dWf = mul(zw, (R)(whole ? C.sqrt_dt : (AR == AR_EXACT ? sqrt(dt) : (dt > 0.0 ? sqrt_pos(dt) : 0.0))));   // :479
        grid_pre(C, dt, dWf);
        tf = tn;                                                         // :482
    }
    __device__ __forceinline__ void grid_pre(const LevelConsts& C, double dt, R dWf) {
        dtc += dt;                                                       // :478
        dWc = add(dWc, dWf);                                             // :480
        if (FN == FN_AVG) {                                              // :554-556
            Gf = add(Gf, half_area(Sf, dt));
            Sf = add(Sf, dx(C, Sf, dt, dWf));
            Gf = add(Gf, half_area(Sf, dt));
        } else {
            Sf = add(Sf, dx(C, Sf, dt, dWf));                            // :481
            if (FN == FN_MIN) Gf = min_<R>(Gf, Sf);                      // :637
        }
    }

    // coarse path catches up at a coarse grid time :483-487
    __device__ __forceinline__ void coarse(const LevelConsts& C, double tn) {
        if (FN == FN_AVG) {                                              // :560-562
            Gc = add(Gc, half_area(Sc, dtc));
            Sc = add(Sc, dx(C, Sc, dtc, dWc));
            Gc = add(Gc, half_area(Sc, dtc));
        } else {
            Sc = add(Sc, dx(C, Sc, dtc, dWc));                           // :484
            if (FN == FN_MIN) Gc = min_<R>(Gc, Sc);                      // :640
        }
        tc = tn; dtc = 0.0; dWc = (R)0;                                  // :485-487
    }

    __device__ __forceinline__ void outputs(const LevelConsts& C, double& vf, double& vc, double& gf, double& gc) const {
        if (FN == FN_AVG) {
            if (AR == AR_EXACT) { vf = __ddiv_rn((double)Gf, C.T); vc = __ddiv_rn((double)Gc, C.T); }   // :571
            else { vf = (double)Gf * C.inv_T; vc = (double)Gc * C.inv_T; }   // production arithmetic: no fp64 division per path
            gf = gc = 0.0;
        } else {
            vf = (double)Sf; vc = (double)Sc; gf = (double)Gf; gc = (double)Gc;
        }
    }
};

// Jump-stream draws for jump number jk of a path: inter-arrival time, jump-size normal and the
// Brownian normal of the adaptive step that ends at that jump.
// (the two halves separately: the batch walk draws the arrival times of a whole warp first and the normals only
// for arrivals that fall inside the path; same words, same values)
// fp64 sampler: Exp(lam) from 52 random bits  (:1294)
__device__ __forceinline__ double gap_f64(const LevelConsts& C, uint32_t hi, uint32_t lo) {
    return neg2_log_u(C, u52(C, hi, lo)) * (0.5 * C.inv_lam);
}
template <int SAMP>
__device__ __forceinline__ double jump_gap(const LevelConsts& C, ull pid, uint32_t jk) {
    if (SAMP == SAMP_F32BM) {
        const U4 r = philox_at(C, pid, STREAM_JUMP, jk);
        const float u = fminf(fmaf(__uint2float_rn(r.x), 0x1p-32f, 0x1p-33f), 0x1.fffffep-1f);
        return (double)(lg2_approx(u) * -0.693147180559945309f) * C.inv_lam;
    } else {
        // two inter-arrival times per block of the gap stream: (x, y) for even jk, (z, w) for odd jk
        const U4 r = philox_at(C, pid, STREAM_GAP, jk >> 1);
        return gap_f64(C, (jk & 1u) ? r.z : r.x, (jk & 1u) ? r.w : r.y);
    }
}
template <typename R, int SAMP>
__device__ __forceinline__ void jump_normals(const LevelConsts& C, ull pid, uint32_t jk, R& zq, R& zw) {
    if (SAMP == SAMP_F32BM) {
        const U4 r = philox_at(C, pid, STREAM_JUMP, jk);
        float a, b;
        bm_pair_f32(C.poly_q4, C.poly_c4, r.z, r.w, a, b);
        zq = (R)a; zw = (R)b;
    } else {
        const U4 s = philox_at(C, pid, STREAM_JUMP, jk);
        double a, b;
        bm_pair_f64<false>(C, s.x, s.y, s.z, a, b);
        zq = (R)a; zw = (R)b;
    }
}
// Jump-stream draws for jump number jk of a path: inter-arrival time, jump-size normal and the
// Brownian normal of the adaptive step that ends at that jump.
template <typename R, int SAMP>
__device__ __forceinline__ void jump_draws(const LevelConsts& C, ull pid, uint32_t jk, double& gap, R& zq, R& zw) {
    if (SAMP == SAMP_F32BM) {
        U4 r = philox_at(C, pid, STREAM_JUMP, jk);
        // Exp(lam) inter-arrival time :454/:474.  Like the normals of this sampler it is an fp32 draw:
        // u = (x + 1/2) 2^-32 from one word, -ln u by MUFU.LG2 (relative error ~1e-7), u kept below 1
        // so that a gap is never zero.
        const float u = fminf(fmaf(__uint2float_rn(r.x), 0x1p-32f, 0x1p-33f), 0x1.fffffep-1f);
        gap = (double)(lg2_approx(u) * -0.693147180559945309f) * C.inv_lam;
        float a, b;
        bm_pair_f32(C.poly_q4, C.poly_c4, r.z, r.w, a, b);
        zq = (R)a; zw = (R)b;
    } else {
        gap = jump_gap<SAMP>(C, pid, jk);
        jump_normals<R, SAMP>(C, pid, jk, zq, zw);
    }
}

// the fine step (1-based, integer-valued double) an arrival time falls into: first j with t < j*dt (:457);
// nsteps+1 for an arrival at or after the end of the path
__device__ __forceinline__ double jump_step_of(const LevelConsts& C, double t) {
    const double nd = (double)C.nsteps;
    double j = nd + 1.0;
    if (t < nd * C.dt) {
        j = fmin(floor(t * C.inv_blk_dt * 4.0) + 1.0, nd);
#pragma unroll 1
        while (j > 1.0 && t < (j - 1.0) * C.dt) j -= 1.0;
#pragma unroll 1
        while (!(t < j * C.dt)) j += 1.0;
    }
    return j;
}

// Merton coupled path, Philox mode.  A 4-step block with no jump inside is exactly a GBM block
// with drift c = r - lam*J_bar; such blocks run in a tight loop identical to the GBM one, up to
// the block that holds the next jump (an integer compare per block).  A block that contains a
// jump runs the jump-adapted steps (rare: lam*T*4/nsteps per path).
//
// Jump records.  The lanes of a warp reach their jumps in different blocks, one lane at a time,
// so whatever a jump costs inside the block is paid 32 times per warp.  Nearly all of it - the
// Philox call, the fp64 log of the inter-arrival time, the Box-Muller pair, the fp64 exp of the
// jump size, the square roots of the two partial steps either side of the jump, and the search
// for the fine step the jump time falls into - does not depend on the path state.  So the first
// kJumpRecs records of every path are produced at the path start with all lanes active and
// parked in shared memory (one column per thread); the divergent block only applies them (a
// dozen FMAs).  A path with more jumps than that (P = 1.9 % at lam*T = 1) produces the further
// ones in place with the same code, reusing the slots.  Values are those of the reference's
// control flow (:455-487): a jump belongs to the first step j with tau < j*dt, its partial step
// starts at the previous jump if that fell in the same step, else at (j-1)*dt.
constexpr int kJumpRecs = MLMCB_JUMP_RECS;          // power of two
enum { REC_D1 = 0, REC_DW1, REC_DJ, REC_D2, REC_SQ2, REC_STEP, REC_FIELDS };

// Where a walker keeps its records.  RingSlots: a ring of kJumpRecs per thread (one column of the block's store),
// refilled in place.  BatchSlots: the compact per-warp store of the batch walk - records 0..n-1 of this path were
// produced by the whole warp before the walk (slots base..base+n-1), anything beyond is produced in place into one
// spare slot (base+n), one at a time.
struct RingSlots {
    double* rec;        // rec[(REC_FIELDS*slot + field)*kThreads]
    __device__ __forceinline__ double& at(uint32_t k, int field) const {
        return rec[(REC_FIELDS * (k & (kJumpRecs - 1)) + field) * kThreads];
    }
};
constexpr int kBatchKmax = 8;           // records per path produced in the batch phase
constexpr int kBatchSlots = 128;        // record slots per warp, one spare per lane included
struct BatchSlots {
    double* rec;        // rec[field*kBatchSlots + slot], already offset to this lane's first slot
    uint32_t n;         // records produced in the batch phase
    __device__ __forceinline__ double& at(uint32_t k, int field) const {
        return rec[field * kBatchSlots + (k < n ? k : n)];
    }
};

template <typename R, int FN, int SAMP, int SCH, typename Slots = RingSlots>
struct MjdWalker {
    const LevelConsts& C;
    ull pid;
    Slots sl;
    JD<R, FN, AR_FAST, SCH> st;
    uint32_t jk;        // index of the next jump of this path
    uint32_t jp;        // records produced so far
    double jnext;       // the fine step (1-based, integer-valued) that jump falls into; nsteps+1: after the end
    double tlast;       // time and step of the newest record produced
    double jlast;
    int cm;
    bool more;          // further records may follow the newest one (always true for the ring store)

    __device__ __forceinline__ double& slot(uint32_t k, int field) const { return sl.at(k, field); }
    // the next record from the jump stream
    __device__ __forceinline__ void produce() {
        double gap;
        R zq, zw;
        jump_draws<R, SAMP>(C, pid, jp, gap, zq, zw);
        const double t = tlast + gap;                                    // :454 / :474
        const double j = jump_step_of(C, t);
        const double d1 = t - (j == jlast ? tlast : (j - 1.0) * C.dt);   // :458
        const double d2 = j * C.dt - t;                                  // :477, if no further jump in step j
        slot(jp, REC_D1) = d1;
        slot(jp, REC_DW1) = (double)(zw * (R)(d1 > 0.0 ? sqrt_pos(d1) : 0.0));   // :460
        slot(jp, REC_DJ) = (double)JD<R, FN, AR_FAST, SCH>::jump_size(C, zq);    // :462
        slot(jp, REC_D2) = d2;
        slot(jp, REC_SQ2) = d2 > 0.0 ? sqrt_pos(d2) : 0.0;               // :479
        slot(jp, REC_STEP) = j;
        tlast = t; jlast = j;
        ++jp;
    }
    __device__ __forceinline__ void start() {
        st.init(C);
        jk = jp = 0; cm = 0;
        tlast = 0.0; jlast = 0.0;
        more = true;
        const double nd = (double)C.nsteps;
#pragma unroll 1
        while (jp < (uint32_t)kJumpRecs) {
            produce();
            if (__all_sync(__activemask(), jlast > nd)) break;           // nobody will look further
        }
        jnext = slot(0, REC_STEP);
    }
    // batch store: records 0..n-1 are in place; t_last / j_last are the arrival time and step of record n-1
    __device__ __forceinline__ void start_batch(uint32_t n, bool more_, double t_last, double j_last) {
        st.init(C);
        jk = 0; jp = n; cm = 0;
        tlast = t_last; jlast = j_last;
        more = more_;
        if (n == 0 && more) produce();                                   // (only when the batch phase was cut short)
        jnext = jk < jp ? slot(0, REC_STEP) : (double)C.nsteps + 1.0;
    }
    // fine step j (:457-482): its jumps, if it has any - only the lanes that do run that loop - then ONE step to the
    // grid time for every lane, over the whole fine step or over what is left of it after the last jump
    __device__ __forceinline__ void fine_step(double j, R z) {
        double h = C.dt, sq = C.sqrt_dt;
        while (jnext == j) {
            st.jump_pre(C, slot(jk, REC_D1), (R)slot(jk, REC_DW1), (R)slot(jk, REC_DJ));
            h = slot(jk, REC_D2); sq = slot(jk, REC_SQ2);
            ++jk;
            if (jk == jp && more) produce();                             // record jk-1 was the newest one
            jnext = jk < jp ? slot(jk, REC_STEP) : (double)C.nsteps + 1.0;
        }
        st.grid_pre(C, h, z * (R)sq);
    }

    // steps j0+1 .. j0+cnt  (:455-487), any M
    template <int MT>
    __device__ __forceinline__ void steps(double j0, int cnt, R z0, R z1, R z2, R z3) {
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const R zk = (k & 2) ? ((k & 1) ? z3 : z2) : ((k & 1) ? z1 : z0);
            fine_step(j0 + (double)(k + 1), zk);                         // :456
            if (MT == 2) { if (k & 1) st.coarse(C, 0.0); }
            else if (MT == 4) { if (k == 3) st.coarse(C, 0.0); }
            else if (++cm == C.M) { cm = 0; st.coarse(C, 0.0); }         // :483
        }
    }
    // steps j0+1 .. j0+4 with the coarse steps at their compile-time places (M = 2 or 4)
    template <int MT>
    __device__ __forceinline__ void jump_block4(double j0, const R z[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            fine_step(j0 + (double)(k + 1), z[k]);
            if (MT == 2 ? (k & 1) : k == 3) st.coarse(C, 0.0);
        }
    }
    // number of leading whole blocks [0, jb) with no jump inside, <= cap
    __device__ __forceinline__ uint32_t jump_block(uint32_t cap) const {
        const double q = (jnext - 1.0) * 0.25;
        return q >= (double)cap ? cap : (uint32_t)q;
    }
};

// One jump-free 4-step block of the Merton path (= a GBM block with drift r - lam*J_bar).
// G accumulates the trapezoid integral (FN_AVG) or the running minimum (FN_MIN).
template <typename R, int FN, int MT, int SCH>
__device__ __forceinline__ void mjd_fast_block(const LevelConsts& C, const StepCoef<R, SCH>& co, const R z[4],
                                               R& Xf, R& Xc, R& Gf, R& Gc) {
    if (FN == FN_AVG) {
        R acc = (R)0.5 * Xf, accc = (R)0.5 * Xc, sz = (R)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            R t_ = co.t(z[k]);
            Xf = fma_<R>(Xf, t_, Xf);
            sz += z[k];
            acc += (k == 3) ? (R)0.5 * Xf : Xf;
            if ((MT == 4 && k == 3) || (MT == 2 && (k & 1))) {
                R tc_ = co.tc(sz);
                Xc = fma_<R>(Xc, tc_, Xc);
                sz = (R)0;
                if (MT == 4) accc = fma_<R>((R)0.5, Xc, accc);
                else accc += (k == 3) ? (R)0.5 * Xc : Xc;
            }
        }
        Gf = fma_<R>((R)C.dt, acc, Gf);
        Gc = fma_<R>((R)C.Mdt, accc, Gc);
    } else if (MT == 4) {
        const R s4 = (z[0] + z[1]) + (z[2] + z[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            R t_ = co.t(z[k]);
            Xf = fma_<R>(Xf, t_, Xf);
            if (FN == FN_MIN) Gf = min_<R>(Gf, Xf);
        }
        R tc_ = co.tc(s4);
        Xc = fma_<R>(Xc, tc_, Xc);
        if (FN == FN_MIN) Gc = min_<R>(Gc, Xc);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            R t0 = co.t(z[k]), t1 = co.t(z[k + 1]);
            Xf = fma_<R>(Xf, t0, Xf);
            if (FN == FN_MIN) Gf = min_<R>(Gf, Xf);
            Xf = fma_<R>(Xf, t1, Xf);
            if (FN == FN_MIN) Gf = min_<R>(Gf, Xf);
            R tc_ = co.tc(z[k] + z[k + 1]);
            Xc = fma_<R>(Xc, tc_, Xc);
            if (FN == FN_MIN) Gc = min_<R>(Gc, Xc);
        }
    }
}

template <typename R, int FN, int MT, int SAMP, int SCH, typename Walker>
__device__ __forceinline__ void mjd_walk(const LevelConsts& C, ull pid, Walker& w,
                                         double& vf, double& vc, double& gf, double& gc);

template <typename R, int FN, int MT, int SAMP, int SCH>
__device__ __forceinline__ void mjd_path_philox(const LevelConsts& C, ull pid, double* rec,
                                                double& vf, double& vc, double& gf, double& gc) {
    MjdWalker<R, FN, SAMP, SCH> w{C, pid, RingSlots{rec}};
    w.start();
    mjd_walk<R, FN, MT, SAMP, SCH>(C, pid, w, vf, vc, gf, gc);
}

// the block-uniform walk of one path whose walker has its first records in place
template <typename R, int FN, int MT, int SAMP, int SCH, typename Walker>
__device__ __forceinline__ void mjd_walk(const LevelConsts& C, ull pid, Walker& w,
                                         double& vf, double& vc, double& gf, double& gc) {
    PhiloxSource<R, SAMP> src{C, pid};
    const StepCoef<R, SCH> co(C);
    const uint32_t nfull = (uint32_t)(C.nsteps >> 2);
    // jb = first block that may hold this path's next jump.  The block loop stays uniform across
    // the warp (all lanes at the same b).  The lanes agree (one REDUX) on how many blocks ahead are
    // jump-free for every one of them and run those in a tight loop identical to the GBM one; the
    // block where some lane jumps is executed once with both paths (divergent), then they agree again.
    uint32_t jb = MT != 0 ? w.jump_block(nfull) : 0;
    R Xf = w.st.Sf, Xc = w.st.Sc, Gf = w.st.Gf, Gc = w.st.Gc;
    uint32_t b = 0;
    while (b < nfull) {
        const unsigned lanes = __activemask();
        const uint32_t run = __reduce_min_sync(lanes, jb > b ? jb - b : 0u);
        const uint32_t stop = b + run;
#pragma unroll kUnrollMjd
        for (; b < stop; ++b) {
            R z[4];
            src.load4(b, z);
            mjd_fast_block<R, FN, MT, SCH>(C, co, z, Xf, Xc, Gf, Gc);
        }
        if (b < nfull) {
            R z[4];
            src.load4(b, z);
            if (b < jb) {
                mjd_fast_block<R, FN, MT, SCH>(C, co, z, Xf, Xc, Gf, Gc);
            } else {
                // the block that holds a jump (or every block when M is generic)
                w.st.Sf = Xf; w.st.Sc = Xc; w.st.Gf = Gf; w.st.Gc = Gc;
                if (MT != 0) w.template jump_block4<MT>((double)(4ull * b), z);
                else w.template steps<MT>((double)(4ull * b), 4, z[0], z[1], z[2], z[3]);
                Xf = w.st.Sf; Xc = w.st.Sc; Gf = w.st.Gf; Gc = w.st.Gc;
                if (MT != 0) jb = w.jump_block(nfull);
            }
            ++b;
        }
    }
    w.st.Sf = Xf; w.st.Sc = Xc; w.st.Gf = Gf; w.st.Gc = Gc;
    const int rem = (int)(C.nsteps & 3);
    if (rem) {                           // l==0 (1 step), M=2,l=1 (2 steps), odd M
        R z[4];
        src.load4(nfull, z);
        w.template steps<MT>((double)(4ull * nfull), rem, z[0], z[1], z[2], z[3]);
    }
    w.st.outputs(C, vf, vc, gf, gc);
}

// ------------------------------------------------------------------------------------------
// Merton, event-driven walk (levels of 2 .. MLMCB_MJD_FLAT_BELOW-1 fine steps; l = 0 has its own loop
// in the kernel below).
// In the block-uniform walk above every lane of a warp is at the same fine step, so a block
// that holds a jump of one lane is run for that lane alone: per warp and path, lam*T*32 serial
// passes through the jump code.  That is noise against 4096 steps but most of the time at
// 4..256 steps - where MLMC puts nearly all of its Merton samples.  Here the lanes are let go
// of each other: every lane owns a queue of paths (the same ids, in the same order, as in the
// grid-stride loop, so the sums are unchanged) and a position in its current path, and the
// warp alternates between two phases that are each uniform in code:
//   advance: every lane walks toward its own next event - jump-free 4-step blocks through the
//            GBM-like block code, single steps inside the block that holds the event - for at
//            most C.run_blocks iterations;
//   event:   every lane that has reached its event handles it: apply the pending jump record
//            (or finish the path, emit it and start the next one), then draw the next record.
// The expensive part of a jump (Philox, log, Box-Muller, exp, square roots, step search) is thus
// paid once per warp and event phase with nearly all lanes taking part, instead of once per lane.
// A lane that reaches its event early idles until the phase ends; run_blocks bounds that wait.
// ------------------------------------------------------------------------------------------
template <typename R, int FN, int MT, int SAMP, int SCH>
struct MjdFlat {
    const LevelConsts& C;
    JD<R, FN, AR_FAST, SCH> st;
    ull pid;
    ull s;              // fine steps completed
    ull ev;             // steps to complete before the event: (step of the next jump) - 1, or nsteps at the path end
    uint32_t jp;        // jump records drawn for this path
    double tlast;       // time and step (1-based; nsteps+1: after the end) of the pending record
    ull jlast;
    double d1, d2, sq2; // the pending record: partial steps either side of the jump ...
    R dW1, dJ;          // ... Brownian increment of the first one, jump size
    R z[4];             // grid normals of block zb
    uint32_t zb;
    int cm;

    __device__ __forceinline__ void produce() {
        double gap;
        R zq, zw;
        jump_draws<R, SAMP>(C, pid, jp, gap, zq, zw);
        ++jp;
        const double t = tlast + gap;                                    // :454 / :474
        const double nd = (double)C.nsteps;
        double j = nd + 1.0;
        if (t < nd * C.dt) {                                             // first j with t < j*dt  (:457)
            j = fmin(floor(t * C.inv_blk_dt * 4.0) + 1.0, nd);
#pragma unroll 1
            while (j > 1.0 && t < (j - 1.0) * C.dt) j -= 1.0;
#pragma unroll 1
            while (!(t < j * C.dt)) j += 1.0;
        }
        const ull ji = (ull)j;
        d1 = t - (ji == jlast ? tlast : (j - 1.0) * C.dt);               // :458
        d2 = j * C.dt - t;                                               // :477, if no further jump in step j
        dW1 = zw * (R)(d1 > 0.0 ? sqrt_pos(d1) : 0.0);                   // :460
        dJ = JD<R, FN, AR_FAST, SCH>::jump_size(C, zq);                  // :462
        sq2 = d2 > 0.0 ? sqrt_pos(d2) : 0.0;                             // :479
        tlast = t; jlast = ji;
        ev = ji - 1;                                                     // ji == nsteps+1 -> the path end
    }
    __device__ __forceinline__ void begin(ull id) {
        pid = id;
        st.init(C);
        s = 0; jp = 0; cm = 0; zb = 0xffffffffu;
        tlast = 0.0; jlast = 0;
    }
    __device__ __forceinline__ void coarse_after_step() {                // s already counts the step
        if (MT == 2) { if (!(s & 1)) st.coarse(C, 0.0); }
        else if (MT == 4) { if (!(s & 3)) st.coarse(C, 0.0); }
        else if (++cm == C.M) { cm = 0; st.coarse(C, 0.0); }             // :483
    }
    __device__ __forceinline__ R zsel() const {
        const int k = (int)(s & 3);
        return (k & 2) ? ((k & 1) ? z[3] : z[2]) : ((k & 1) ? z[1] : z[0]);
    }
    // walk toward the event; on return the normals of the block that holds step s+1 are loaded
    __device__ __forceinline__ void advance(const StepCoef<R, SCH>& co) {
        uint32_t budget = C.run_blocks;
        for (;;) {
            const uint32_t blk = (uint32_t)(s >> 2);
            if (zb != blk && s < C.nsteps) {
                PhiloxSource<R, SAMP>{C, pid}.load4(blk, z);
                zb = blk;
            }
            if (!(s < ev && budget)) break;
            --budget;
            if (MT != 0 && !(s & 3) && s + 4 <= ev) {
                mjd_fast_block<R, FN, MT, SCH>(C, co, z, st.Sf, st.Sc, st.Gf, st.Gc);
                s += 4;
            } else {
                st.grid_pre(C, C.dt, zsel() * (R)C.sqrt_dt);
                ++s;
                coarse_after_step();
            }
        }
    }
};

__device__ __forceinline__ unsigned char* mjd_shared_bytes();
#ifndef MLMCB_MJD_BATCH_FROM
#define MLMCB_MJD_BATCH_FROM 8      // fp64 sampler: levels from this many fine steps take the batch walk, shorter ones the event-driven
#endif                              // walk (profiles/r02_merton_levels.txt); the fp32 sampler's cheap records favour the event-driven walk
// which of the two short-path walks a level takes (host and device agree through this one function)
__host__ __device__ __forceinline__ bool mjd_use_batch(int samp, ull nsteps) {
    return samp == SAMP_F64BM && nsteps >= (ull)MLMCB_MJD_BATCH_FROM;
}
template <typename R, int FN, int MT, int SAMP, int SCH>
__device__ __forceinline__ void mjd_batch_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                               double* Pf_out, double* Pc_out);

// Body of the event-driven Merton kernel for block `bid` of the `nblk` blocks that share this level's paths
// (a whole grid for a level call, a slice of the grid inside a sweep).
template <typename R, int FN, int MT, int SAMP, int SCH>
__device__ __forceinline__ void mjd_flat_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                              double* Pf_out, double* Pc_out) {
    const ull stride = (ull)nblk * blockDim.x;
    ull i = (ull)bid * blockDim.x + threadIdx.x;
    if (C.nsteps == 1) {
        // l = 0: a path is its jumps and one closing step, every one of them "draw a record, step to
        // min(jump time, T)".  Each lane runs that loop over its own queue of paths, so the warp never
        // waits for the lane with the most jumps.  The closing step takes the Brownian normal of the
        // record that overshoots T (that record's jump never happens, and its normal is independent of
        // its arrival time), so a path of N jumps costs exactly N+1 Philox calls.
        bool active = i < C.n_paths;
        JD<R, FN, AR_FAST, SCH> st;
        st.init(C);
        double t = 0.0;
        uint32_t k = 0;
        const double tend = C.dt;                                        // (double)1 * dt, the grid time  :456
        while (__any_sync(0xffffffffu, active)) {
            if (active) {
                double gap;
                R zq, zw;
                jump_draws<R, SAMP>(C, C.path_offset + i, k, gap, zq, zw);
                const double tau = t + gap;                              // :454 / :474
                const bool jump = tau < tend;                            // :457
                const double d = (jump ? tau : tend) - t;                // :458 / :477
                const R dW = zw * (R)(d > 0.0 ? sqrt_pos(d) : 0.0);      // :460 / :479
                // the diffusion sub-step is the same whether it ends at a jump or at T; the jump factor is evaluated on
                // every lane (a lane that closes its path discards it): cheaper than running the two cases one after
                // the other on half a warp each
                const R u = st.Sf + st.dx(C, st.Sf, d, dW);              // :465 / :481
                if (FN == FN_AVG) st.Gf = st.Gf + st.half_area(st.Sf, d) + st.half_area(u, d);   // :532-534 / :554-556
                const R J = (R)exp_bounded(C, fma(C.js, (double)zq, C.jm));                       // 1 + dJ  :462
                st.Sf = jump ? u * J : u;
                if (FN == FN_MIN) st.Gf = min_<R>(st.Gf, st.Sf);         // :621 / :637
                if (jump) {
                    // the coarse path of a one-step level takes the very increments of the fine one at a jump and does
                    // not move otherwise: it IS the fine state after the jump (:468, :622)
                    st.Sc = st.Sf; st.Gc = st.Gf;
                    t = tau; ++k;
                } else {
                    double vf, vc, gf, gc;
                    st.outputs(C, vf, vc, gf, gc);
                    emit_t<FN>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
                    i += stride;
                    active = i < C.n_paths;
                    st.init(C);
                    t = 0.0; k = 0;
                }
            }
        }
        return;
    }
    bool active = i < C.n_paths;
    MjdFlat<R, FN, MT, SAMP, SCH> w{C};
    const StepCoef<R, SCH> co(C);
    if (active) { w.begin(C.path_offset + i); w.produce(); }
    while (__any_sync(0xffffffffu, active)) {
        if (active) w.advance(co);
        if (active && w.s == w.ev) {
            bool draw = true;
            double d2o = 0.0, sq2o = 0.0;
            const bool jump = w.jlast <= C.nsteps;
            if (jump) {                                                  // the jump(s) of step s+1  (:457-474)
                w.st.jump_pre(C, w.d1, w.dW1, w.dJ);
                d2o = w.d2; sq2o = w.sq2;
            } else {                                                     // path end
                double vf, vc, gf, gc;
                w.st.outputs(C, vf, vc, gf, gc);
                emit_t<FN>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
                i += stride;
                active = draw = i < C.n_paths;
                if (active) w.begin(C.path_offset + i);
            }
            if (draw) w.produce();
            if (jump && w.jlast != w.s + 1) {                            // no further jump in this step: finish it  (:477-482)
                w.st.grid_pre(C, d2o, w.zsel() * (R)sq2o);
                ++w.s;
                w.coarse_after_step();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Merton, batch walk (levels of 2 .. MLMCB_MJD_FLAT_BELOW-1 fine steps - where mlmc puts its Merton samples).
// A warp takes 32 paths at a time (lane = path) and goes through three phases, each uniform in code:
//   arrivals  every lane draws inter-arrival times until its next arrival falls outside the path (or kBatchKmax are
//             drawn): Philox + log only, the arrival times parked in shared memory.  Rounds = the largest count in the
//             warp plus one; the cheap part of a record is all that runs at partial occupancy.
//   records   the arrivals of the WHOLE warp (lam*T*32 of them on average) are dealt out one per lane - a prefix sum
//             gives every lane its slots, a work list maps slot -> (owner lane, k) - and the expensive part of a record
//             (second Philox block, Box-Muller pair, exp of the jump size, two square roots, step search) is computed
//             with nearly all lanes busy, whoever owns the path, and stored in slot order.
//   walk      the block-uniform walk of the long-path kernel (mjd_walk): all lanes at the same 4-step block, grid
//             normals drawn with all lanes active, jump-free blocks through the GBM-like block code, blocks that hold
//             a jump through the jump-adapted steps, reading the records a lane owns from its slots.
// A path with more than kBatchKmax jumps, or a warp whose arrivals do not fit kBatchSlots, produces the excess
// records in place inside the walk (one spare slot per lane), exactly as the ring store does.  Same draws, same
// arithmetic and the same fast-block / jump-block rule as the other walks: per-path payoffs are bit-identical.
// ------------------------------------------------------------------------------------------
struct BatchStore {
    double t[kBatchKmax][32];                   // arrival times, t[k][lane]
    double rec[REC_FIELDS * kBatchSlots];       // records, rec[field*kBatchSlots + slot]
    unsigned short work[kBatchSlots];           // work list of the record phase: owner lane | k << 5
};

constexpr size_t kFwStoreBytes = 8 * 32 * 8 + 6 * MLMCB_FW_SLOTS * 8 + MLMCB_FW_SLOTS * 4 + MLMCB_FW_SLOTS * 2;    // sizeof(FwStore<FN_AVG>), the factor walk's store (below)
constexpr size_t kMjdSharedWarp = sizeof(BatchStore) > kFwStoreBytes ? sizeof(BatchStore) : kFwStoreBytes;
constexpr size_t kMjdSharedBytes = kMjdSharedWarp * (kThreads / 32) > sizeof(double) * kJumpRecs * REC_FIELDS * kThreads
                                       ? kMjdSharedWarp * (kThreads / 32) : sizeof(double) * kJumpRecs * REC_FIELDS * kThreads;
__device__ __forceinline__ unsigned char* mjd_shared_bytes() {
    __shared__ __align__(16) unsigned char bytes[kMjdSharedBytes];
    return bytes;
}

template <typename R, int FN, int MT, int SAMP, int SCH>
__device__ __forceinline__ void mjd_batch_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                               double* Pf_out, double* Pc_out) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    BatchStore& bs = reinterpret_cast<BatchStore*>(mjd_shared_bytes())[warp];
    const unsigned FULL = 0xffffffffu;
    const double tend = (double)C.nsteps * C.dt;
    const ull nbatch = (C.n_paths + 31) >> 5;
    for (ull batch = (ull)bid * wpb + warp; batch < nbatch; batch += (ull)nblk * wpb) {
        const ull i = (batch << 5) + lane;
        const bool valid = i < C.n_paths;
        const ull pid = C.path_offset + i;
        // ---- arrivals ----
        double t = 0.0;
        uint32_t n = 0;
        bool more = valid;                      // the next arrival may still fall inside the path
        while (__any_sync(FULL, more && n < (uint32_t)kBatchKmax)) {
            if (more && n < (uint32_t)kBatchKmax) {
                t += jump_gap<SAMP>(C, pid, n);                          // :454 / :474
                if (t < tend) { bs.t[n][lane] = t; ++n; }
                else more = false;
            }
        }
        // ---- slots: n records + one spare per lane; if the warp's arrivals do not fit, one record per lane ----
        uint32_t total = __reduce_add_sync(FULL, n + 1u);
        if (total > (uint32_t)kBatchSlots) {
            if (n > 1u) { n = 1u; more = true; }
            total = __reduce_add_sync(FULL, n + 1u);
        }
        uint32_t base = n + 1u, wbase = n;      // inclusive scans -> exclusive below
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL, base, o), b = __shfl_up_sync(FULL, wbase, o);
            if (lane >= (unsigned)o) { base += a; wbase += b; }
        }
        const uint32_t nwork = __shfl_sync(FULL, wbase, 31);
        base -= n + 1u; wbase -= n;
        for (uint32_t k = 0; k < n; ++k) bs.work[wbase + k] = (unsigned short)(lane | (k << 5));
        __syncwarp();
        // ---- records: one arrival per lane, whoever owns the path ----
        for (uint32_t w0 = 0; w0 < nwork; w0 += 32) {
            const uint32_t wi = w0 + lane;
            const bool act = wi < nwork;
            const uint32_t item = act ? bs.work[wi] : 0u, ol = item & 31u, k = item >> 5;
            const uint32_t obase = __shfl_sync(FULL, base, ol);
            if (act) {
                const double tk = bs.t[k][ol], tp = k ? bs.t[k - 1][ol] : 0.0;
                R zq, zw;
                jump_normals<R, SAMP>(C, C.path_offset + (batch << 5) + ol, k, zq, zw);
                const double j = jump_step_of(C, tk), jprev = k ? jump_step_of(C, tp) : 0.0;
                const double d1 = tk - (j == jprev ? tp : (j - 1.0) * C.dt);             // :458
                const double d2 = j * C.dt - tk;                                         // :477
                double* r = bs.rec + obase + k;
                r[REC_D1 * kBatchSlots] = d1;
                r[REC_DW1 * kBatchSlots] = (double)(zw * (R)(d1 > 0.0 ? sqrt_pos(d1) : 0.0));       // :460
                r[REC_DJ * kBatchSlots] = (double)JD<R, FN, AR_FAST, SCH>::jump_size(C, zq);        // :462
                r[REC_D2 * kBatchSlots] = d2;
                r[REC_SQ2 * kBatchSlots] = d2 > 0.0 ? sqrt_pos(d2) : 0.0;                           // :479
                r[REC_STEP * kBatchSlots] = j;
            }
        }
        __syncwarp();
        // ---- walk ----
        MjdWalker<R, FN, SAMP, SCH, BatchSlots> w{C, pid, BatchSlots{bs.rec + base, n}};
        const double t_last = n ? bs.t[n - 1][lane] : 0.0;
        w.start_batch(n, more, t_last, n ? w.slot(n - 1, REC_STEP) : 0.0);
        double vf, vc, gf, gc;
        mjd_walk<R, FN, MT, SAMP, SCH>(C, pid, w, vf, vc, gf, gc);
        if (valid) emit_t<FN>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
        __syncwarp();                           // the store is rewritten by the next batch
    }
}

template <typename R, int FN, int MT, int SAMP, int SCH>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? MLMCB_MINBLOCKS_F64_MJD : MLMCB_MINBLOCKS_FLAT)
mjd_batch_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
                 double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    mjd_batch_body<R, FN, MT, SAMP, SCH>(C, blockIdx.x, gridDim.x, acc, Pf_out, Pc_out);
    grid_finish(acc, partials, ticket, out7);
}

template <typename R, int FN, int MT, int SAMP, int SCH>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? MLMCB_MINBLOCKS_F64_MJD : MLMCB_MINBLOCKS_FLAT)
mjd_flat_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
                double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    mjd_flat_body<R, FN, MT, SAMP, SCH>(C, blockIdx.x, gridDim.x, acc, Pf_out, Pc_out);
    grid_finish(acc, partials, ticket, out7);
}

// ------------------------------------------------------------------------------------------
// Merton, factor walk (Euler-Maruyama, fp64 paths, M = 2 or 4, every level of 2 .. 2^30 fine steps).
//
// An Euler-Maruyama step of the Merton SDE multiplies the state by a factor that does not depend on the state:
//   grid step          S <- S (1 + t),            t  = mu h + sig sqrt(h) z
//   step to a jump     S <- S (1 + t1) J,         t1 = mu d1 + sig sqrt(d1) zw,   J = exp(jm + js zq) = 1 + dJ   (:462-465)
// and the coarse increment over any stretch of fine sub-steps is the SUM of their t's (both are linear in (h, dW)),
// so the coarse path needs one accumulator: S_c <- S_c (1 + sum t) at every coarse grid time and at every jump.
// Everything about a jump is therefore known before the walk starts.  The closing step of a fine step that holds a
// jump - from the jump to the grid time, d2 long, driven by the grid normal z_j of that step - is affine in the factor
// the sampler delivers for the whole step: with t = a + b z_j,
//   t' = mu d2 + sig sqrt(d2) z_j = kappa + rho t,   rho = sqrt(d2/dt),  kappa = mu d2 - rho a.
// A record is (t1, J, rho, kappa, step) [+ d1, d2 for the trapezoid of the Asian payoff]; it is produced with all lanes
// busy (arrivals of the whole warp dealt out one per lane, as in the batch walk) and APPLYING it is four multiply-adds.
// The walk itself is uniform across the warp: every lane is at the same fine step; octets (eight steps, three Philox
// blocks) in which no lane has a jump run the GBM octet code, the others run step by step - "lanes with a jump in this
// step apply it, then every lane closes the step with t' = fma(rho, t, kappa)" (rho = 1, kappa = 0 without a jump).
// No lane ever waits for another lane's jump arithmetic: what is left of the divergence is four FMAs per jump.
//
// Rounds.  Records are produced for up to kFwKmax arrivals per path and round into kFwSlots slots per warp (prefix
// sum, n + 1 slots per lane: the last one is a sentinel).  A path with more arrivals, or a warp whose arrivals do not
// fit (then 3 per lane), is walked as far as the records of EVERY lane are complete - all steps below the step of the
// last record of a lane that may have more - the jumps already known for the next step are applied, and the next
// round starts from each lane's last applied record (records produced but not applied are produced again: the draws
// are a pure function of (path, k)).  Any lam*T terminates: the lane that limits the walk applies all of its records.
// Same draws and the same per-path control flow as the other walks (:449-491); values differ from theirs by rounding
// only (products of factors instead of increments: ~1e-15 relative).
// ------------------------------------------------------------------------------------------
constexpr int kFwKmax = 8;              // arrivals per path and round
constexpr int kFwSlots = MLMCB_FW_SLOTS; // record slots per warp, sentinels included
constexpr uint32_t kFwCap = kFwSlots / 32 - 1;   // records per lane and round that always fit (with the sentinels)
enum { FW_T1 = 0, FW_J, FW_RHO, FW_KAPPA, FW_D1, FW_D2 };
template <int FN>
struct FwStore {
    static constexpr int kFields = FN == FN_AVG ? 6 : 4;
    double t[kFwKmax][32];              // arrival times of this round, t[k][lane]
    double rec[kFields][kFwSlots];
    int step[kFwSlots];                 // fine step (1-based) the record falls into; sentinel: INT_MAX
    unsigned short work[kFwSlots];      // work list of the record phase: owner lane | k << 5
};
static_assert(sizeof(FwStore<FN_AVG>) <= kFwStoreBytes, "kFwStoreBytes");
constexpr ull kFwMaxSteps = 1ull << 30; // step numbers are 32-bit here; longer paths take the block-uniform walk
// which levels take the factor walk (host and device agree through this one function); flag bits as in mlmcb.h:
// 0x10 = MLMCB_FP32_PATHS, 0x40 = MLMCB_MILSTEIN
__host__ __device__ __forceinline__ bool mjd_use_factor(int flags, int mt, ull nsteps) {
    return !(flags & (0x10 | 0x40)) && mt != 0 && nsteps >= 2 && nsteps <= kFwMaxSteps;
}

// step factors t_i = a + b z_i of octet o (steps 8o+1 .. 8o+8), or of its first NP pairs
template <int SAMP>
struct FactorSource {
    const LevelConsts& C;
    ull pid;
    __device__ __forceinline__ void load8(uint32_t o, double t[8]) const {
        if (SAMP == SAMP_F64BM) {
            PhiloxSource<double, SAMP>{C, pid}.load8_affine(o, t, C.a_f, C.b_f);
        } else {
            float f[4], g[4];
            bm_quad_f32(C, philox_at(C, pid, STREAM_GRID, 2u * o), f);
            bm_quad_f32(C, philox_at(C, pid, STREAM_GRID, 2u * o + 1u), g);
#pragma unroll
            for (int k = 0; k < 4; ++k) { t[k] = fma(C.b_f, (double)f[k], C.a_f); t[4 + k] = fma(C.b_f, (double)g[k], C.a_f); }
        }
    }
    // paths of 2 or 4 fine steps: the first one or two pairs of octet 0
    __device__ __forceinline__ void load_head(int npairs, double t[8]) const {
        if (SAMP == SAMP_F64BM) {
            const U4 r = philox_at(C, pid, STREAM_GRID, 0u);
            bm_pair_f64<true>(C, r.x, r.y, r.z, t[0], t[1], C.a_f, C.b_f);
            if (npairs > 1) {
                const U4 s = philox_at(C, pid, STREAM_GRID, 1u);
                bm_pair_f64<true>(C, r.w, s.x, s.y, t[2], t[3], C.a_f, C.b_f);
            }
        } else {
            float f[4];
            bm_quad_f32(C, philox_at(C, pid, STREAM_GRID, 0u), f);
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = fma(C.b_f, (double)f[k], C.a_f);
        }
    }
};

template <int FN, int MT, int SAMP>
struct FactorWalk {
    const LevelConsts& C;
    FwStore<FN>& bs;
    double Sf, Sc, Gf, Gc;      // G: trapezoid integral (FN_AVG) / running minimum (FN_MIN)
    double tsum;                // sum of the fine t's since the last coarse step = the coarse t
    double dtc, h;              // FN_AVG: time since the last coarse step; length of the closing step of the current fine step
    double rho, kappa;          // closing step of the current fine step: t' = fma(rho, t, kappa)
    uint32_t jk;                // slot of this lane's next record
    int jnext;                  // its fine step

    __device__ __forceinline__ void init() {
        Sf = Sc = C.X0;
        Gf = Gc = FN == FN_MIN ? C.X0 : 0.0;
        tsum = 0.0; dtc = 0.0; h = C.dt; rho = 1.0; kappa = 0.0;
    }
    // this lane's next record: the sub-step to the jump and the jump itself, on both paths (:458-474)
    __device__ __forceinline__ void apply_record() {
        const double t1 = bs.rec[FW_T1][jk], J = bs.rec[FW_J][jk];
        rho = bs.rec[FW_RHO][jk]; kappa = bs.rec[FW_KAPPA][jk];
        tsum += t1;
        if (FN == FN_AVG) {                                              // :532-540
            const double d1 = bs.rec[FN == FN_AVG ? FW_D1 : 0][jk];
            h = bs.rec[FN == FN_AVG ? FW_D2 : 0][jk];
            dtc += d1;
            const double uf = fma(Sf, t1, Sf), uc = fma(Sc, tsum, Sc);
            Gf = fma(0.5 * d1, Sf + uf, Gf);
            Gc = fma(0.5 * dtc, Sc + uc, Gc);
            Sf = uf * J; Sc = uc * J;
            dtc = 0.0;
        } else {                                                         // :465-468, :621-622
            Sf = fma(Sf, t1, Sf) * J;
            Sc = fma(Sc, tsum, Sc) * J;
            if (FN == FN_MIN) { Gf = min_<double>(Gf, Sf); Gc = min_<double>(Gc, Sc); }
        }
        tsum = 0.0;
        ++jk;
        jnext = bs.step[jk];
    }
    // the jumps of fine step j known so far (:457-474), lanes that have one only; every lane of the warp calls it
    __device__ __forceinline__ void jumps(int j) {
        while (__any_sync(0xffffffffu, jnext == j)) {
            if (jnext == j) apply_record();
        }
    }
    // the step to the grid time that closes fine step j (:477-487); t = a + b z_j
    template <bool COARSE_END>
    __device__ __forceinline__ void close(double t) {
        const double tt = fma(rho, t, kappa);
        if (FN == FN_AVG) {                                              // :554-556
            const double uf = fma(Sf, tt, Sf);
            Gf = fma(0.5 * h, Sf + uf, Gf);
            Sf = uf;
            dtc += h;
            h = C.dt;
        } else {
            Sf = fma(Sf, tt, Sf);                                        // :481
            if (FN == FN_MIN) Gf = min_<double>(Gf, Sf);                 // :637
        }
        tsum += tt;
        rho = 1.0; kappa = 0.0;
        if (COARSE_END) {                                                // :483-487
            if (FN == FN_AVG) {
                const double uc = fma(Sc, tsum, Sc);
                Gc = fma(0.5 * dtc, Sc + uc, Gc);
                Sc = uc;
                dtc = 0.0;
            } else {
                Sc = fma(Sc, tsum, Sc);
                if (FN == FN_MIN) Gc = min_<double>(Gc, Sc);
            }
            tsum = 0.0;
        }
    }
    // fine step j: its jumps, then the closing step; every lane of the warp calls it
    template <bool COARSE_END>
    __device__ __forceinline__ void step(int j, double t) {
        jumps(j);
        close<COARSE_END>(t);
    }
    // A whole octet that every lane may close and in which SOME lanes have jumps: per coarse interval, the lanes that
    // have a jump in it walk its MT steps on their own (the very operations of step(), so a path's value does not depend
    // on which of the two forms ran) and turn their factors of the interval into no-ops; then every lane runs the
    // jump-free interval.  A lane without jumps never waits for more than the few lanes' MT steps.
    __device__ __forceinline__ void mixed_octet(int j0, double t[8]) {
        const double mdt = MT == 4 ? ((C.dt + C.dt) + C.dt) + C.dt : C.dt + C.dt;
#pragma unroll
        for (int g = 0; g < 8; g += MT) {
            double tc = MT == 4 ? ((t[g] + t[g + 1]) + t[g + 2]) + t[g + 3] : t[g] + t[g + 1];
            double hdt = 0.5 * C.dt, hmdt = 0.5 * mdt;                   // FN_AVG: zero on a lane that walked the interval itself
            if (__any_sync(0xffffffffu, jnext <= j0 + g + MT)) {
                if (jnext <= j0 + g + MT) {
#pragma unroll
                    for (int k = 0; k < MT; ++k) {
                        while (jnext == j0 + g + k + 1) apply_record();
                        if (k == MT - 1) close<true>(t[g + k]);
                        else close<false>(t[g + k]);
                        t[g + k] = 0.0;
                    }
                    tc = 0.0; hdt = 0.0; hmdt = 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < MT; ++k) {
                if (FN == FN_AVG) {
                    const double uf = fma(Sf, t[g + k], Sf);
                    Gf = fma(hdt, Sf + uf, Gf);
                    Sf = uf;
                } else {
                    Sf = fma(Sf, t[g + k], Sf);
                    if (FN == FN_MIN) Gf = min_<double>(Gf, Sf);
                }
            }
            if (FN == FN_AVG) {
                const double uc = fma(Sc, tc, Sc);
                Gc = fma(hmdt, Sc + uc, Gc);
                Sc = uc;
            } else {
                Sc = fma(Sc, tc, Sc);
                if (FN == FN_MIN) Gc = min_<double>(Gc, Sc);
            }
        }
    }
    // an octet in which no lane has a jump (tsum = 0, dtc = 0 on entry and on exit).  The arithmetic is that of step()
    // without a jump, operation for operation - sequential sums for the coarse factor and the coarse step length, the
    // trapezoid step by step - so that a path's value does not depend on what the other lanes of its warp do.
    __device__ __forceinline__ void free_octet(const double t[8]) {
        const double hdt = 0.5 * C.dt;
        const double mdt = MT == 4 ? ((C.dt + C.dt) + C.dt) + C.dt : C.dt + C.dt;
#pragma unroll
        for (int g = 0; g < 8; g += MT) {
            const double tc = MT == 4 ? ((t[g] + t[g + 1]) + t[g + 2]) + t[g + 3] : t[g] + t[g + 1];
#pragma unroll
            for (int k = 0; k < MT; ++k) {
                if (FN == FN_AVG) {
                    const double uf = fma(Sf, t[g + k], Sf);
                    Gf = fma(hdt, Sf + uf, Gf);
                    Sf = uf;
                } else {
                    Sf = fma(Sf, t[g + k], Sf);
                    if (FN == FN_MIN) Gf = min_<double>(Gf, Sf);
                }
            }
            if (FN == FN_AVG) {
                const double uc = fma(Sc, tc, Sc);
                Gc = fma(0.5 * mdt, Sc + uc, Gc);
                Sc = uc;
            } else {
                Sc = fma(Sc, tc, Sc);
                if (FN == FN_MIN) Gc = min_<double>(Gc, Sc);
            }
        }
    }
    __device__ __forceinline__ void outputs(double& vf, double& vc, double& gf, double& gc) const {
        if (FN == FN_AVG) { vf = Gf * C.inv_T; vc = Gc * C.inv_T; gf = gc = 0.0; }       // :571
        else { vf = Sf; vc = Sc; gf = Gf; gc = Gc; }
    }
};

template <int FN, int MT, int SAMP>
__device__ __forceinline__ void mjd_factor_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                                double* Pf_out, double* Pc_out) {
    static_assert(MT == 2 || MT == 4, "the factor walk is built for M = 2 and M = 4");
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    FwStore<FN>& bs = reinterpret_cast<FwStore<FN>*>(mjd_shared_bytes())[warp];
    const unsigned FULL = 0xffffffffu;
    const int nsteps = (int)C.nsteps;
    const double tend = (double)C.nsteps * C.dt;
    const ull nbatch = (C.n_paths + 31) >> 5;
    for (ull batch = (ull)bid * wpb + warp; batch < nbatch; batch += (ull)nblk * wpb) {
        const ull i = (batch << 5) + lane;
        const bool valid = i < C.n_paths;
        const ull pid0 = C.path_offset + (batch << 5), pid = pid0 + lane;
        FactorWalk<FN, MT, SAMP> w{C, bs};
        w.init();
        const FactorSource<SAMP> src{C, pid};
        uint32_t k_done = 0;            // records applied in earlier rounds
        double t_done = 0.0;            // arrival time and fine step of the last of them
        int j_done = 0;
        bool more = valid;              // the next arrival may still fall inside the path
        int j_cur = 1;                  // the fine step to close next (the same in every lane)
        bool mid = false;               // jumps of step j_cur were applied before its records were complete (end of a round)
        for (;;) {
            // ---- arrivals: inter-arrival times until the path ends or kFwKmax are drawn (:454, :474) ----
            double t = t_done;
            uint32_t n = 0;
            while (__any_sync(FULL, more && n < (uint32_t)kFwKmax)) {
                if (more && n < (uint32_t)kFwKmax) {
                    if (SAMP == SAMP_F64BM) {                            // one block of the gap stream holds two
                        const uint32_t idx = k_done + n;
                        const U4 r = philox_at(C, pid, STREAM_GAP, idx >> 1);
                        if (!(idx & 1u)) {
                            t += gap_f64(C, r.x, r.y);
                            if (t < tend) { bs.t[n][lane] = t; ++n; }
                            else more = false;
                        }
                        if (more && n < (uint32_t)kFwKmax) {
                            t += gap_f64(C, r.z, r.w);
                            if (t < tend) { bs.t[n][lane] = t; ++n; }
                            else more = false;
                        }
                    } else {
                        t += jump_gap<SAMP>(C, pid, k_done + n);
                        if (t < tend) { bs.t[n][lane] = t; ++n; }
                        else more = false;
                    }
                }
            }
            // ---- slots: n records and a sentinel per lane; three records per lane always fit ----
            if (__reduce_add_sync(FULL, n + 1u) > (uint32_t)kFwSlots && n > kFwCap) { n = kFwCap; more = true; }
            uint32_t base = n + 1u, wbase = n;  // inclusive scans -> exclusive below
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(FULL, base, o), b = __shfl_up_sync(FULL, wbase, o);
                if (lane >= (unsigned)o) { base += a; wbase += b; }
            }
            const uint32_t nwork = __shfl_sync(FULL, wbase, 31);
            base -= n + 1u; wbase -= n;
            for (uint32_t k = 0; k < n; ++k) bs.work[wbase + k] = (unsigned short)(lane | (k << 5));
            bs.step[base + n] = 0x7fffffff;
            __syncwarp();
            // ---- records: one arrival per lane, whoever owns the path ----
            for (uint32_t w0 = 0; w0 < nwork; w0 += 32) {
                const uint32_t wi = w0 + lane;
                const bool act = wi < nwork;
                const uint32_t item = act ? bs.work[wi] : 0u, ol = item & 31u, k = item >> 5;
                const uint32_t obase = __shfl_sync(FULL, base, ol), okd = __shfl_sync(FULL, k_done, ol);
                const double otd = __shfl_sync(FULL, t_done, ol);
                const int ojd = __shfl_sync(FULL, j_done, ol);
                const uint32_t s = obase + k;
                double tk = 0.0;
                int j = 0;
                if (act) {
                    tk = bs.t[k][ol];
                    j = (int)jump_step_of(C, tk);
                    bs.step[s] = j;
                }
                __syncwarp();                   // the step of the previous arrival of the same path: slot s - 1
                if (act) {
                    const double tp = k ? bs.t[k - 1][ol] : otd;
                    const int jprev = k ? bs.step[s - 1] : ojd;
                    double zq, zw;
                    jump_normals<double, SAMP>(C, pid0 + ol, okd + k, zq, zw);
                    const double d1 = tk - (j == jprev ? tp : (double)(j - 1) * C.dt);      // :458
                    const double d2 = (double)j * C.dt - tk;                                // :477
                    const double dW1 = zw * (d1 > 0.0 ? sqrt_pos(d1) : 0.0);                // :460
                    const double r = (d2 > 0.0 ? sqrt_pos(d2) : 0.0) * C.inv_sqrt_dt;       // :479
                    bs.rec[FW_T1][s] = fma(C.sig, dW1, C.mu * d1);
                    bs.rec[FW_J][s] = exp_bounded(C, fma(C.js, zq, C.jm));                  // :462
                    bs.rec[FW_RHO][s] = r;
                    bs.rec[FW_KAPPA][s] = fma(-r, C.a_f, C.mu * d2);
                    if (FN == FN_AVG) { bs.rec[FN == FN_AVG ? FW_D1 : 0][s] = d1; bs.rec[FN == FN_AVG ? FW_D2 : 0][s] = d2; }
                }
            }
            __syncwarp();
            w.jk = base;
            w.jnext = bs.step[base];
            // steps 1 .. hz can be closed: every lane's records for them are complete
            const int hz = (int)__reduce_min_sync(FULL, (unsigned)(more && n ? bs.step[base + n - 1] - 1 : nsteps));
            // ---- walk ----
            while (j_cur <= hz) {
                const uint32_t o = (uint32_t)(j_cur - 1) >> 3;
                if (nsteps >= 8 && ((j_cur - 1) & 7) == 0 && !mid) {
                    // whole octets ahead in which no lane has a jump, as far as they may be closed
                    const int lim = w.jnext <= hz ? w.jnext - 1 : hz;                       // steps 1 .. lim are jump-free in this lane
                    const uint32_t stop = __reduce_min_sync(FULL, (unsigned)lim >> 3);      // octets [o, stop)
                    if (stop > o) {
#pragma unroll kUnrollFw
                        for (uint32_t q = o; q < stop; ++q) {
                            double tq[8];
                            src.load8(q, tq);
                            w.free_octet(tq);
                        }
                        j_cur = (int)(stop << 3) + 1;
                        continue;
                    }
                }
                double tq[8];
                if (nsteps >= 8) {
                    src.load8(o, tq);
                    // a whole closable octet with jumps in some lanes.  (Final-value products only: with the trapezoid or
                    // the running minimum in the 128-register budget the extra code costs the short levels 1-2 % and
                    // gains 3-5 % from 128 steps on - measured, not kept.)
                    if constexpr (FN == FN_FINAL) {
                        if (nsteps >= MLMCB_FW_MIXED_FROM && ((j_cur - 1) & 7) == 0 && !mid && (int)(o << 3) + 8 <= hz) {
                            w.mixed_octet((int)(o << 3), tq);
                            j_cur += 8;
                            continue;
                        }
                    }
                } else {
                    src.load_head(nsteps >> 1, tq);
                    if (hz == nsteps && j_cur == 1) {                   // the whole path in one go (2 or 4 steps)
                        if (MT == 2) {
                            w.template step<false>(1, tq[0]);
                            w.template step<true>(2, tq[1]);
                            if (nsteps == 4) { w.template step<false>(3, tq[2]); w.template step<true>(4, tq[3]); }
                        } else {
                            w.template step<false>(1, tq[0]);
                            w.template step<false>(2, tq[1]);
                            w.template step<false>(3, tq[2]);
                            w.template step<true>(4, tq[3]);
                        }
                        j_cur = nsteps + 1;
                        break;
                    }
                }
                const int j0 = (int)(o << 3);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = j0 + k + 1;
                    if (j >= j_cur && j <= hz) {
                        if ((k % MT) == MT - 1) w.template step<true>(j, tq[k]);
                        else w.template step<false>(j, tq[k]);
                    }
                }
                j_cur = j0 + 8 < hz ? j0 + 9 : hz + 1;
                mid = false;
            }
            if (j_cur > nsteps) break;
            // ---- the walk waits for records: apply what is known of step j_cur, restart from the last applied record ----
            w.jumps(j_cur);
            mid = true;
            const uint32_t c = w.jk - base;
            if (c) { t_done = bs.t[c - 1][lane]; j_done = bs.step[base + c - 1]; k_done += c; }
            if (c < n) more = true;
            __syncwarp();
        }
        double vf, vc, gf, gc;
        w.outputs(vf, vc, gf, gc);
        if (valid) emit_t<FN, false>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
        __syncwarp();                           // the store is rewritten by the next batch
    }
}

template <int FN, int MT, int SAMP>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? MLMCB_MINBLOCKS_F64_MJD : MLMCB_MINBLOCKS_FLAT)
mjd_factor_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
                  double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    mjd_factor_body<FN, MT, SAMP>(C, blockIdx.x, gridDim.x, acc, Pf_out, Pc_out);
    grid_finish(acc, partials, ticket, out7);
}

// ------------------------------------------------------------------------------------------
// American put under GBM: Amer_GBM.path :1511-1615 + Amer_payoff :138-158 (M == 2 only, :1524),
// with the demo notebook's exercise boundary b(t) = max(K*(1-c*sqrt(T-t)), 0) (cell 1, `exb`).
// Everything that depends only on the step index - barrier values, exercise values, discount
// factors - is the same for every path, so a small set-up kernel tabulates it once per launch
// (one 64-byte row per fine step, evaluated in the reference's operation order); the path kernel
// then reads one row per step (same address across the warp -> broadcast) and keeps the three
// state components of the fine and of the coarse path in registers.
// The reference's `halfdWc=dWc` (:1577) aliases the array that `dWc += dWf` (:1549) updates in
// place, so at the coarse step halfdWc IS dWc; the mid-point (:1585) is reproduced as such.
// ------------------------------------------------------------------------------------------
struct AmerRow {
    double bR;        // b(j*dt): right barrier of fine step j, left barrier of step j+1      :1561
    double mkF, eF;   // max(K - b(tfMid), 0) and exp(-r*tfMid), tfMid = (j-1)*dt + dt/2     :1558-1569
    double mkC, eC;   // j even: max(K - b(tcMid), 0), exp(-r*tcMid), tcMid = (j-M)*dt + M*dt/2   :1590-1605
    double mb1;       // j even: 0.5*(b(j*dt) + b((j-M)*dt))                                   :1599
    double pad0, pad1;
};

__device__ __forceinline__ double amer_eb(const LevelConsts& C, double t) {
    return fmax(__dmul_rn(C.K, __dsub_rn(1.0, __dmul_rn(C.boundary_c, __dsqrt_rn(__dsub_rn(C.T, t))))), 0.0);
}

#ifndef MLMCB_TEMPLATES_ONLY   // the one non-template kernel: defined in mlmcb_api.cu only
__global__ void amer_table_kernel(const __grid_constant__ LevelConsts C, AmerRow* tab) {
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > C.nsteps) return;
    AmerRow row = {0, 0, 0, 0, 0, 0, 0, 0};
    const double tj = __dmul_rn((double)j, C.dt);
    row.bR = amer_eb(C, tj);
    if (j >= 1) {
        const double t_ = __dmul_rn((double)(j - 1), C.dt), tm = __dadd_rn(t_, __ddiv_rn(C.dt, 2.0));
        row.mkF = fmax(__dsub_rn(C.K, amer_eb(C, tm)), 0.0);
        row.eF = exp(__dmul_rn(-C.r, tm));
        if (j % 2 == 0) {
            const double tc = __dmul_rn((double)(j - 2), C.dt), tcm = __dadd_rn(tc, __ddiv_rn(C.Mdt, 2.0));
            row.mkC = fmax(__dsub_rn(C.K, amer_eb(C, tcm)), 0.0);
            row.eC = exp(__dmul_rn(-C.r, tcm));
            row.mb1 = __dmul_rn(0.5, __dadd_rn(row.bR, amer_eb(C, tc)));
        }
    }
    tab[j] = row;
}
#endif

template <int AR>
struct AmerState {
    double f0, f1, f2, c0, c1, c2, dWc, lbF, lbC;

    __device__ __forceinline__ void init(const LevelConsts& C, const AmerRow* tab) {
        f0 = c0 = C.X0; f1 = c1 = 1.0; f2 = c2 = 0.0; dWc = 0.0;
        lbF = lbC = __ldg(&tab[0].bR);
    }
    // probability of touching the barrier inside a step (Brownian-bridge formula) :1566 / :1600
    __device__ __forceinline__ static double touch_exact(const LevelConsts& C, double A, double B, double Xl) {
        return exp(__ddiv_rn(__dmul_rn(__dmul_rn(-2.0, A), B), __dmul_rn(__dmul_rn(C.sig2, C.dt), __dmul_rn(Xl, Xl))));
    }
    // production form: the exponent with a precomputed w = -2/(sig^2 dt Xl^2); far from the barrier
    // the exponent is below -750 for every lane of the warp (exp underflows to exactly 0), and the
    // whole warp skips the exp.
    __device__ __forceinline__ static double inv_scale(const LevelConsts& C, double Xl) {
        const double g = Xl * Xl;
        double y = rcp64_seed(g);
        y = fma(y, fma(-g, y, 1.0), y);
        y = fma(y, fma(-g, y, 1.0), y);
        return C.amer_cneg * y;
    }
    __device__ __forceinline__ static double exp_or_zero(const LevelConsts& C, double arg) {
        if (__all_sync(__activemask(), arg < -750.0)) return 0.0;
        return arg < -700.0 ? 0.0 : exp_bounded(C, arg);
    }
    __device__ __forceinline__ static double inc_exact(const LevelConsts& C, double X, double h, double dW) {
        return __dadd_rn(__dmul_rn(__dmul_rn(C.mu, X), h), __dmul_rn(__dmul_rn(C.sig, X), dW));
    }
    // fine step j with normal z; row = tab[j]
    __device__ __forceinline__ void fine(const LevelConsts& C, const AmerRow& row, double z) {
        double Xl = f0, Xr, dWf;
        if (AR == AR_EXACT) {
            dWf = __dmul_rn(C.sqrt_dt, z);                              // :1548
            dWc = __dadd_rn(dWc, dWf);
            Xr = __dadd_rn(Xl, inc_exact(C, Xl, C.dt, dWf));            // :1554
        } else {
            dWc += z;                                                    // in units of sqrt(dt)
            Xr = fma(Xl, fma(C.b_f, z, C.a_f), Xl);
        }
        const double A = pos_part(Xl - lbF), B = pos_part(Xr - row.bR);
        const double P1 = AR == AR_EXACT ? touch_exact(C, A, B, Xl) : exp_or_zero(C, inv_scale(C, Xl) * A * B);
        if (AR == AR_EXACT) {
            f2 = __dadd_rn(f2, __dmul_rn(__dmul_rn(__dmul_rn(f1, P1), row.mkF), row.eF));      // :1569
            f1 = __dmul_rn(f1, __dsub_rn(1.0, P1));                      // :1571
        } else {
            f2 = fma(f1 * P1, row.mkF * row.eF, f2);
            f1 = fma(-f1, P1, f1);
        }
        f0 = Xr;
        lbF = row.bR;
    }
    // coarse step ending at even j; row = tab[j]
    __device__ __forceinline__ void coarse(const LevelConsts& C, const AmerRow& row) {
        double Xl = c0, Xr, Xmid;
        if (AR == AR_EXACT) {
            Xr = __dadd_rn(Xl, inc_exact(C, Xl, C.Mdt, dWc));            // :1584
            Xmid = __dadd_rn(__dadd_rn(Xl, __dmul_rn(0.5, __dsub_rn(Xr, Xl))),
                             __dmul_rn(__dmul_rn(C.sig, Xl), __dsub_rn(dWc, __dmul_rn(0.5, dWc))));    // :1585
        } else {
            Xr = fma(Xl, fma(C.b_f, dWc, C.a_c), Xl);
            Xmid = fma(C.b_f * Xl, 0.5 * dWc, fma(0.5, Xr - Xl, Xl));
        }
        const double A = pos_part(Xl - lbC), Bm = pos_part(Xmid - row.mb1), B = pos_part(Xr - row.bR);
        double P11, P12;                                                 // :1600-1601
        if (AR == AR_EXACT) {
            P11 = touch_exact(C, A, Bm, Xl); P12 = touch_exact(C, Bm, B, Xl);
        } else {
            const double w = inv_scale(C, Xl) * Bm;
            P11 = exp_or_zero(C, w * A); P12 = exp_or_zero(C, w * B);
        }
        if (AR == AR_EXACT) {
            const double P1 = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, P11), __dsub_rn(1.0, P12)));   // :1603
            c2 = __dadd_rn(c2, __dmul_rn(__dmul_rn(__dmul_rn(c1, P1), row.mkC), row.eC));         // :1605
            c1 = __dmul_rn(c1, __dsub_rn(1.0, P1));
        } else {
            const double keep = (1.0 - P11) * (1.0 - P12);
            c2 = fma(c1 * (1.0 - keep), row.mkC * row.eC, c2);
            c1 *= keep;
        }
        c0 = Xr;
        lbC = row.bR;
        dWc = 0.0;
    }
    // Amer_payoff :153-157
    __device__ __forceinline__ void payoffs(const LevelConsts& C, double& Pf, double& Pc) const {
        Pf = __dadd_rn(f2, __dmul_rn(__dmul_rn(f1, pos_part(__dsub_rn(C.K, f0))), C.disc));
        Pc = C.is_l0 ? 0.0 : __dadd_rn(c2, __dmul_rn(__dmul_rn(c1, pos_part(__dsub_rn(C.K, c0))), C.disc));
    }
};

__device__ __forceinline__ AmerRow load_row(const AmerRow* tab, ull j) {
    const double2* p = reinterpret_cast<const double2*>(tab + j);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    AmerRow r = {a.x, a.y, b.x, b.y, c.x, c.y, 0.0, 0.0};
    return r;
}

template <int SAMP>
__global__ void __launch_bounds__(kThreads)
amer_level_kernel(const __grid_constant__ LevelConsts C, const AmerRow* tab, double* partials, unsigned* ticket,
                  double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        PhiloxSource<double, SAMP> src{C, C.path_offset + i};
        AmerState<AR_FAST> st;
        st.init(C, tab);
        const uint32_t nblk = (uint32_t)(C.nsteps >> 2);
        const AmerRow* rows = tab + 1;                                   // row of fine step j = 4b+k+1
        for (uint32_t b = 0; b < nblk; ++b, rows += 4) {
            double z[4];
            src.load4(b, z);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const AmerRow row = load_row(rows, k);
                st.fine(C, row, z[k]);
                if (k & 1) st.coarse(C, row);
            }
        }
        const int rem = (int)(C.nsteps & 3);                             // l = 0 (1 step), l = 1 (2 steps)
        if (rem) {
            double z[4];
            src.load4(nblk, z);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < rem) {
                    const AmerRow row = load_row(rows, k);
                    st.fine(C, row, z[k]);
                    if (k & 1) st.coarse(C, row);
                }
            }
        }
        double Pf, Pc;
        st.payoffs(C, Pf, Pc);
        if (C.raw) {      // Amer_GBM.path returns Xf[:,0..2], Xc[:,0..2] (:1615)
            const ull n = C.n_paths;
            C.raw[i] = st.f0; C.raw[n + i] = st.f1; C.raw[2 * n + i] = st.f2;
            C.raw[3 * n + i] = st.c0; C.raw[4 * n + i] = st.c1; C.raw[5 * n + i] = st.c2;
        }
        acc.add(Pf, Pc, C.is_l0);
        if (Pf_out) Pf_out[i] = Pf;
        if (Pc_out) Pc_out[i] = Pc;
    }
    grid_finish(acc, partials, ticket, out7);
}

template <int AR>
__global__ void __launch_bounds__(kThreads)
replay_amer_kernel(const __grid_constant__ LevelConsts C, const AmerRow* tab, const double* Z, double* partials,
                   unsigned* ticket, double* out7, double* Pf_out, double* Pc_out) {
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        AmerState<AR> st;
        st.init(C, tab);
        for (ull j = 1; j <= C.nsteps; ++j) {
            const AmerRow row = load_row(tab, j);
            st.fine(C, row, Z[(j - 1) * C.n_paths + i]);
            if ((j & 1) == 0) st.coarse(C, row);
        }
        double Pf, Pc;
        st.payoffs(C, Pf, Pc);
        if (C.raw) {      // Amer_GBM.path returns Xf[:,0..2], Xc[:,0..2] (:1615)
            const ull n = C.n_paths;
            C.raw[i] = st.f0; C.raw[n + i] = st.f1; C.raw[2 * n + i] = st.f2;
            C.raw[3 * n + i] = st.c0; C.raw[4 * n + i] = st.c1; C.raw[5 * n + i] = st.c2;
        }
        acc.add(Pf, Pc, C.is_l0);
        if (Pf_out) Pf_out[i] = Pf;
        if (Pc_out) Pc_out[i] = Pc;
    }
    grid_finish(acc, partials, ticket, out7);
}

// ------------------------------------------------------------------------------------------
// Kernels.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit(const LevelConsts& C, ull i, double vf, double vc, double gf, double gc,
                                     Sums7& acc, double* Pf_out, double* Pc_out) {
    const double Pf = payoff_fine(C, vf, gf);
    const double Pc = C.is_l0 ? vc : payoff_coarse(C, vc, gc);          // l==0: raw coarse output :82
    acc.add(Pf, Pc, C.is_l0);
    if (Pf_out) Pf_out[i] = Pf;
    if (Pc_out) Pc_out[i] = Pc;
    if (C.raw) {      // what the reference path function returns: (Xf,Xc) | (Af/T,Ac/T) | (Xf,Mf,Xc,Mc)
        C.raw[i] = vf; C.raw[C.n_paths + i] = gf; C.raw[2 * C.n_paths + i] = vc; C.raw[3 * C.n_paths + i] = gc;
    }
}

// the block's jump-record store (Merton kernels only): the ring store of the long-path walk and the batch store of
// the short-path walk live in the same shared bytes - a block serves one level, so it only ever uses one of them
template <int MODEL>
struct JumpRecStore {
    __device__ __forceinline__ static double* column() { return nullptr; }
};
template <>
struct JumpRecStore<1> {
    __device__ __forceinline__ static double* column() { return reinterpret_cast<double*>(mjd_shared_bytes()) + threadIdx.x; }
};

// GBM paths of one or two fine steps (l = 0; M = 2, l = 1), a kernel of their own so that the long-path
// kernel's register allocation and code placement do not move when this one is tuned.
// NS = 0: one or two fine steps per path (several paths share a Philox block);  NS = 4 / -4: four fine steps
// with M = 4 (l = 1) / M = 2 (l = 2) - one draw block IS the path, so there is no block loop, the coarse steps
// sit at compile-time places and the payoff is the compile-time one of the path functional.  Same draws as
// the long-path kernel would use (STREAM_GRID block 0 of the path), so results do not depend on which runs.
template <typename R, int FN, int SAMP, int SCH, int NS>
__device__ __forceinline__ void gbm_packed_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                                double* Pf_out, double* Pc_out) {
    const ull stride = (ull)nblk * blockDim.x;
    const ull tid = (ull)bid * blockDim.x + threadIdx.x;
    if (NS != 0) {
        const StepCoef<R, SCH> co(C);
        const R X0 = (R)C.X0;
        const double h0 = 0.5 * C.X0;
        // fp64 sampler: the words of the thread's NEXT path are drawn in the body that turns the current ones into
        // normals (integer work next to FP64 work, see gbm_path_fast PIPE)
        U4 wcur[2] = {}, wnext[2] = {};
        if (SAMP == SAMP_F64BM && MLMCB_PIPE_SHORT && tid < C.n_paths) PhiloxSource<R, SAMP>{C, C.path_offset + tid}.words4(wcur);
        for (ull i = tid; i < C.n_paths; i += stride) {
            R z[4];
            if (SAMP == SAMP_F64BM && MLMCB_PIPE_SHORT) {
                PhiloxSource<R, SAMP>{C, C.path_offset + i + stride}.words4(wnext);      // one path past the end on the last pass: unused
                PhiloxSource<R, SAMP>{C, 0}.normals4(wcur, z);
                wcur[0] = wnext[0]; wcur[1] = wnext[1];
            } else {
                PhiloxSource<R, SAMP>{C, C.path_offset + i}.load4(0, z);
            }
            R Xf = X0, Xc = X0;
            R Ff = FN == FN_MIN ? X0 : (R)0, Fc = Ff;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                Xf = fma_<R>(Xf, co.t(z[k]), Xf);
                if (FN == FN_AVG) Ff += Xf;
                if (FN == FN_MIN) Ff = min_<R>(Ff, Xf);
                if (NS == -4 ? (k & 1) : k == 3) {
                    const R s = NS == -4 ? z[k - 1] + z[k] : (z[0] + z[1]) + (z[2] + z[3]);
                    Xc = fma_<R>(Xc, co.tc(s), Xc);
                    if (FN == FN_AVG) Fc += Xc;
                    if (FN == FN_MIN) Fc = min_<R>(Fc, Xc);
                }
            }
            double vf, vc, gf = 0.0, gc = 0.0;
            if (FN == FN_AVG) {          // trapezoid :369-386
                vf = C.avg_wf * ((double)Ff + fma(-0.5, (double)Xf, h0));
                vc = C.avg_wc * ((double)Fc + fma(-0.5, (double)Xc, h0));
            } else {
                vf = (double)Xf; vc = (double)Xc; gf = (double)Ff; gc = (double)Fc;
            }
            emit_t<FN, false>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
        }
        return;
    }
    // The coarsest levels (1 or 2 fine steps per path) are where MLMC puts most of its samples and
    // a path needs fewer normals than one Philox block holds: 4 (or 2) consecutive global path ids
    // share one block (stream STREAM_PACK, counter = id / paths-per-block), each taking its own
    // normals.  Still a pure function of (seed, level, global path id).
    const int per = C.nsteps == 1 ? 4 : 2;
    const ull first = C.path_offset / per, last = (C.path_offset + C.n_paths - 1) / per;
    const ull lo = C.path_offset, hi = C.path_offset + C.n_paths;
    if (C.nsteps == 1) {
        // l == 0: one Euler step, no coarse path (:341), sums row [Pf, Pf^2, Pf, Pf^2, 0, 0, 0] (:949)
        const StepCoef<R, SCH> co(C);
        const R X0 = (R)C.X0;
        const double h0 = 0.5 * C.X0, vc0 = FN == FN_AVG ? 0.0 : C.X0, gc0 = FN == FN_MIN ? C.X0 : 0.0;
        double s0 = 0.0, s1 = 0.0;
        // one group of four ids through the checked path: bounds test per id, optional per-path outputs
        auto checked_group = [&](ull c) {
            R z[4];
            PhiloxSource<R, SAMP>{C, c}.load4_stream(STREAM_PACK, 0, z);
            const ull p0 = c << 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const R Xf = fma_<R>(X0, co.t(z[k]), X0);
                double vf = (double)Xf, gf = 0.0;
                if (FN == FN_AVG) vf = C.avg_wf * ((double)Xf + fma(-0.5, (double)Xf, h0));      // :369-386
                if (FN == FN_MIN) gf = (double)min_<R>(X0, Xf);                                    // :420
                const double Pf = payoff_fine(C, vf, gf);
                if (p0 + k >= lo && p0 + k < hi) {
                    s0 += Pf;
                    s1 = fma(Pf, Pf, s1);
                    const ull i = p0 + k - lo;
                    if (Pf_out) Pf_out[i] = Pf;
                    if (Pc_out) Pc_out[i] = vc0;                                                   // raw coarse output :82
                    if (C.raw) {
                        C.raw[i] = vf; C.raw[C.n_paths + i] = gf;
                        C.raw[2 * C.n_paths + i] = vc0; C.raw[3 * C.n_paths + i] = gc0;
                    }
                }
            }
        };
        if (Pf_out || Pc_out || C.raw) {
            for (ull c = first + tid; c <= last; c += stride) checked_group(c);
        } else {
            // Sums only - the loop MLMC spends most of its samples in: the groups that lie wholly inside the request
            // (in_first <= c < in_end) run without bounds tests or stores; the at most two ragged groups at the ends
            // of the id range take the checked path on two threads of the grid.
            const ull in_first = (lo + 3) >> 2, in_end = hi >> 2;
            for (ull c = in_first + tid; c < in_end; c += stride) {         // (pipelining this loop across groups measured 2-5 % slower)
                R z[4];
                PhiloxSource<R, SAMP>{C, c}.load4_stream(STREAM_PACK, 0, z);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const R Xf = fma_<R>(X0, co.t(z[k]), X0);
                    double vf = (double)Xf, gf = 0.0;
                    if (FN == FN_AVG) vf = C.avg_wf * ((double)Xf + fma(-0.5, (double)Xf, h0));
                    if (FN == FN_MIN) gf = (double)min_<R>(X0, Xf);
                    const double Pf = payoff_fine(C, vf, gf);
                    s0 += Pf;
                    s1 = fma(Pf, Pf, s1);
                }
            }
            if (tid == 0 && first < in_first) checked_group(first);
            if (tid == 1 % stride && last >= in_end && !(first < in_first && last == first)) checked_group(last);
        }
        acc.s[0] = acc.s[2] = s0;
        acc.s[1] = acc.s[3] = s1;
        acc.s[4] = acc.s[5] = acc.s[6] = 0.0;
    } else {
        for (ull c = first + tid; c <= last; c += stride) {
            PhiloxSource<R, SAMP> quad{C, c};
            PackedSource<R> src;
            quad.load4_stream(STREAM_PACK, 0, src.z);
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const ull pid = c * 2 + k;
                if (pid < lo || pid >= hi) continue;
                src.base = 2 * k;
                double vf, vc, gf, gc;
                gbm_path_fast<R, FN, 0, SCH>(C, src, vf, vc, gf, gc);
                emit_t<FN>(C, pid - lo, vf, vc, gf, gc, acc, Pf_out, Pc_out);
            }
        }
    }
}

template <typename R, int FN, int SAMP, int SCH, int NS>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? MLMCB_MINBLOCKS_F64_SHORT : MLMCB_MINBLOCKS)
gbm_packed_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
                  double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    gbm_packed_body<R, FN, SAMP, SCH, NS>(C, blockIdx.x, gridDim.x, acc, Pf_out, Pc_out);
    grid_finish(acc, partials, ticket, out7);
}


// VAR (fp64-sampler GBM only): 1 = the sweep kernel's build - one draw block in flight, fewer registers;
// 2 = the long-path build - software-pipelined octets (gbm_path_fast PIPE)
template <typename R, int MODEL, int FN, int MT, int SAMP, int SCH, int VAR = 0>
__device__ __forceinline__ void level_body(const LevelConsts& C, unsigned bid, unsigned nblk, Sums7& acc,
                                           double* Pf_out, double* Pc_out) {
    const ull stride = (ull)nblk * blockDim.x;
    double* const rec = JumpRecStore<MODEL>::column();
    for (ull i = (ull)bid * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        const ull pid = C.path_offset + i;
        double vf, vc, gf, gc;
        if (MODEL == 0) {
            PhiloxSource<R, SAMP> src{C, pid};
            gbm_path_fast<R, FN, MT, SCH, PhiloxSource<R, SAMP>, VAR == 1 ? 1 : PhiloxSource<R, SAMP>::kUnroll, VAR == 2>(C, src, vf, vc, gf, gc);
        } else {
            mjd_path_philox<R, FN, MT, SAMP, SCH>(C, pid, rec, vf, vc, gf, gc);
        }
        emit_t<FN>(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
    }
}

template <typename R, int MODEL, int FN, int MT, int SAMP, int SCH, int VAR = 0>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? (MODEL == 0 ? (VAR == 1 ? MLMCB_MINBLOCKS_F64_SHORT : MLMCB_MINBLOCKS_F64) : MLMCB_MINBLOCKS_F64_MJD)
                                                                : (MODEL == 0 ? MLMCB_MINBLOCKS : MLMCB_MINBLOCKS_MJD))
level_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
             double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    level_body<R, MODEL, FN, MT, SAMP, SCH, VAR>(C, blockIdx.x, gridDim.x, acc, Pf_out, Pc_out);
    grid_finish(acc, partials, ticket, out7);
}

// ------------------------------------------------------------------------------------------
// One launch for a whole sweep of Option.mlmc's `for l in range(L+1)` loop (:994-997).
// The grid is cut into consecutive block ranges, one per requested level (deepest first: it has the
// longest serial path and must start first); a block looks its range up, runs the body of the kernel
// that level would get on its own - packed / four-step / long GBM paths, event-driven / block-uniform
// Merton walk - on that level's constants, and leaves one row of partial sums.  The last block to
// arrive folds every level's rows in block order into out[row][8] and clears the ticket, so the
// persistent workspace needs neither a memset nor an allocation between launches.  The per-level
// constants sit in the kernel's parameter block (indexed by a block-uniform level number).
// ------------------------------------------------------------------------------------------
constexpr int kSweepMax = 20;
enum { SK_PACKED = 0, SK_FOUR = 1, SK_LONG = 2, SK_MJD_FLAT = 3, SK_MJD_UNIFORM = 4, SK_MJD_FACTOR = 5 };
struct SweepConsts {
    int n_levels, n_rows;
    unsigned first[kSweepMax + 1];      // blocks [first[k], first[k+1]) simulate entry k
    int kind[kSweepMax], row[kSweepMax];
    LevelConsts lv[kSweepMax];
};
static_assert(sizeof(SweepConsts) <= 32764, "kernel parameters are limited to 32764 bytes");

template <typename R, int MODEL, int FN, int MT, int SAMP, int SCH>
__global__ void __launch_bounds__(kThreads, (SAMP != SAMP_F32BM) ? (MODEL == 0 ? MLMCB_MINBLOCKS_F64_SHORT : MLMCB_MINBLOCKS_F64_MJD)
                                                                : (MODEL == 0 ? MLMCB_MINBLOCKS - 2 : MLMCB_MINBLOCKS_FLAT - 1))
sweep_kernel(const __grid_constant__ SweepConsts S, double* partials, unsigned* ticket, double* out) {
    sampler_init<SAMP>();
    int k = 0;
    while (k + 1 < S.n_levels && blockIdx.x >= S.first[k + 1]) ++k;
    const LevelConsts& C = S.lv[k];
    const unsigned bid = blockIdx.x - S.first[k], nblk = S.first[k + 1] - S.first[k];
    Sums7 acc;
    acc.clear();
    if (MODEL == 0) {
        const int kind = S.kind[k];
        if (kind == SK_PACKED) gbm_packed_body<R, FN, SAMP, SCH, 0>(C, bid, nblk, acc, nullptr, nullptr);
        else if (MT != 0 && kind == SK_FOUR) gbm_packed_body<R, FN, SAMP, SCH, MT == 4 ? 4 : -4>(C, bid, nblk, acc, nullptr, nullptr);
        else level_body<R, 0, FN, MT, SAMP, SCH, 1>(C, bid, nblk, acc, nullptr, nullptr);
    } else {
        if (S.kind[k] == SK_MJD_FACTOR) {
            if constexpr (sizeof(R) == 8 && SCH == SCH_EM && MT != 0) mjd_factor_body<FN, MT, SAMP>(C, bid, nblk, acc, nullptr, nullptr);
        } else if (S.kind[k] == SK_MJD_FLAT) {
            if (mjd_use_batch(SAMP, C.nsteps)) mjd_batch_body<R, FN, MT, SAMP, SCH>(C, bid, nblk, acc, nullptr, nullptr);
            else mjd_flat_body<R, FN, MT, SAMP, SCH>(C, bid, nblk, acc, nullptr, nullptr);
        }
        else level_body<R, 1, FN, MT, SAMP, SCH>(C, bid, nblk, acc, nullptr, nullptr);
    }
    __shared__ double sm[(kThreads / 32) * 8];
    __shared__ unsigned last;
    const double mine = block_fold(acc, sm);
    if (threadIdx.x < 7) partials[(size_t)blockIdx.x * 8 + threadIdx.x] = mine;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int t = threadIdx.x; t < S.n_rows * 8; t += blockDim.x) out[t] = 0.0;      // rows of empty requests
    for (int j = 0; j < S.n_levels; ++j) {
        Sums7 t;
        t.clear();
        for (unsigned r = S.first[j] + threadIdx.x; r < S.first[j + 1]; r += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 7; ++q) t.s[q] += __ldcg(&partials[(size_t)r * 8 + q]);
        }
        const double tot = block_fold(t, sm);
        if (threadIdx.x < 7) out[S.row[j] * 8 + threadIdx.x] = tot;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

template <int FN, int AR, int SCH>
__global__ void __launch_bounds__(kThreads)
replay_gbm_kernel(const __grid_constant__ LevelConsts C, const double* Z, double* partials, unsigned* ticket,
                  double* out7, double* Pf_out, double* Pc_out) {
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        double vf, vc, gf, gc;
        if (AR == AR_EXACT) {
            gbm_path_exact<FN, SCH>(C, Z, C.n_paths, i, vf, vc, gf, gc);
        } else {
            ReplaySource src{Z, C.n_paths, i, C.nsteps};
            gbm_path_fast<double, FN, 0, SCH>(C, src, vf, vc, gf, gc);
        }
        emit(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
    }
    grid_finish(acc, partials, ticket, out7);
}

template <typename R, int FN, int MT, int SAMP, int SCH>
__global__ void __launch_bounds__(kThreads)
level_anti_kernel(const __grid_constant__ LevelConsts C, double* partials, unsigned* ticket,
                  double* out7, double* Pf_out, double* Pc_out) {
    sampler_init<SAMP>();
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        PhiloxSource<R, SAMP> src{C, C.path_offset + i};
        double vf, va, vc;
        gbm_anti_path_fast<R, FN, MT, SCH>(C, src, vf, va, vc);
        emit_anti(C, i, vf, va, vc, acc, Pf_out, Pc_out);
    }
    grid_finish(acc, partials, ticket, out7);
}

template <int FN, int AR, int SCH>
__global__ void __launch_bounds__(kThreads)
replay_gbm_anti_kernel(const __grid_constant__ LevelConsts C, const double* Z, double* partials, unsigned* ticket,
                       double* out7, double* Pf_out, double* Pc_out) {
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        double vf, va, vc;
        if (AR == AR_EXACT) {
            gbm_anti_path_exact<FN, SCH>(C, Z, C.n_paths, i, vf, va, vc);
        } else {
            ReplaySource src{Z, C.n_paths, i, C.nsteps};
            gbm_anti_path_fast<double, FN, 0, SCH>(C, src, vf, va, vc);
        }
        emit_anti(C, i, vf, va, vc, acc, Pf_out, Pc_out);
    }
    grid_finish(acc, partials, ticket, out7);
}

template <int FN, int AR, int SCH>
__global__ void __launch_bounds__(kThreads)
replay_mjd_kernel(const __grid_constant__ LevelConsts C, const double* zW, const long long* offW,
                  const double* zQ, const long long* offQ, const double* E, const long long* offE,
                  double* partials, unsigned* ticket, double* out7, double* Pf_out, double* Pc_out) {
    Sums7 acc;
    acc.clear();
    const ull stride = (ull)gridDim.x * blockDim.x;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < C.n_paths; i += stride) {
        JD<double, FN, AR, SCH> st;
        st.init(C);
        long long iw = offW[i], iq = offQ[i], ie = offE[i];
        double tau = 0.0;
        tau += E[ie++];                                                  // :454
        int cm = 0;
        for (ull j = 1; j <= C.nsteps; ++j) {
            // :456  tn = j*T/Nsteps   (EXACT keeps the multiply-then-divide)
            const double tn = AR == AR_EXACT ? __ddiv_rn(__dmul_rn((double)j, C.T), (double)C.nsteps)
                                             : (double)j * C.dt;
            while (tau < tn) {
                st.jump(C, tau, zW[iw++], zQ[iq++]);
                tau += E[ie++];                                          // :474
            }
            st.grid(C, tn, zW[iw++]);
            if (++cm == C.M) { cm = 0; st.coarse(C, tn); }
        }
        double vf, vc, gf, gc;
        st.outputs(C, vf, vc, gf, gc);
        emit(C, i, vf, vc, gf, gc, acc, Pf_out, Pc_out);
    }
    grid_finish(acc, partials, ticket, out7);
}

// The jump-stream records of the Merton kernels: for record k of a path the inter-arrival time (already
// divided by lam), the jump-size normal and the Brownian normal of the sub-step that ends at the jump.
template <int SAMP>
__global__ void jump_probe_kernel(const __grid_constant__ LevelConsts C, ull n_records, double* out) {
    sampler_init<SAMP>();
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_paths) return;
    for (ull k = 0; k < n_records; ++k) {
        double gap, zq, zw;
        jump_draws<double, SAMP>(C, C.path_offset + i, (uint32_t)k, gap, zq, zw);
        out[(0 * n_records + k) * C.n_paths + i] = gap;
        out[(1 * n_records + k) * C.n_paths + i] = zq;
        out[(2 * n_records + k) * C.n_paths + i] = zw;
    }
}

// The normals of the coarsest levels (1 or 2 fine steps per path), where 4 or 2 neighbouring global
// path ids share one Philox block of the STREAM_PACK stream: out[step*n_paths + path].
template <int SAMP>
__global__ void packed_probe_kernel(const __grid_constant__ LevelConsts C, int nsteps, double* out) {
    sampler_init<SAMP>();
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_paths) return;
    const int per = nsteps == 1 ? 4 : 2;
    const ull pid = C.path_offset + i;
    double z[4];
    PhiloxSource<double, SAMP>{C, pid / per}.load4_stream(STREAM_PACK, 0, z);
    const int base = (int)(pid % per) * nsteps;
    for (int s = 0; s < nsteps; ++s) {
        const int k = base + s;
        out[(ull)s * C.n_paths + i] = k == 0 ? z[0] : k == 1 ? z[1] : k == 2 ? z[2] : z[3];
    }
}

template <int SAMP>
__global__ void sampler_probe_kernel(const __grid_constant__ LevelConsts C, ull n_steps, double* out) {
    sampler_init<SAMP>();
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C.n_paths) return;
    PhiloxSource<double, SAMP> src{C, C.path_offset + i};
    for (uint32_t b = 0; 4 * (ull)b < n_steps; ++b) {
        double z[4];
        src.load4(b, z);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * (ull)b + k < n_steps) out[(4 * (ull)b + k) * C.n_paths + i] = z[k];
    }
}

}  // namespace mlmcb
