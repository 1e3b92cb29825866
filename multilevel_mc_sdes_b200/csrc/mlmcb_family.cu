// libmlmcb - the level kernels of ONE family and ONE sampler (see mlmcb_pick.cuh, mlmcb_kernels.cuh):
//   MLMCB_TU_FAMILY 0  GBM: level_kernel<R, 0, ...>, gbm_packed_kernel, level_anti_kernel
//                   1  Merton, block-uniform walk: level_kernel<R, 1, ...>
//                   2  Merton, short paths: mjd_flat_kernel (event-driven walk, l = 0 loop), mjd_batch_kernel (batch walk)
//                   3  GBM, one launch per sweep:  sweep_kernel<double, 0, ...>
//                   4  Merton, one launch per sweep: sweep_kernel<double, 1, ...>
//                   5  Merton, factor walk: mjd_factor_kernel (Euler-Maruyama, fp64 paths, M = 2 | 4, 2 .. 2^30 fine steps)
//   MLMCB_TU_SAMP   0  f32bm, 1 f64bm
#define MLMCB_TEMPLATES_ONLY
#include "mlmcb_pick.cuh"

#if !defined(MLMCB_TU_FAMILY) || !defined(MLMCB_TU_SAMP)
#error "compile with -DMLMCB_TU_FAMILY=0..5 -DMLMCB_TU_SAMP=0|1 (multilevel_mc_sdes_b200/build.py does)"
#endif

namespace mlmcb {
namespace {
constexpr int S = MLMCB_TU_SAMP;
#if MLMCB_TU_FAMILY == 0
struct Gbm {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return level_kernel<R, 0, FN, MT, SAMP, SCH>; }
};
template <int NS>
struct GbmPacked {      // one, two or four fine steps per path: the coarseness M is fixed by the step count
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return gbm_packed_kernel<R, FN, SAMP, SCH, NS>; }
};
struct GbmLong {        // fp64 sampler, Euler-Maruyama, fp64 paths, M = 2 | 4: software-pipelined octets
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() {
        if constexpr (SAMP == SAMP_F64BM && sizeof(R) == 8 && SCH == SCH_EM && MT != 0) return level_kernel<R, 0, FN, MT, SAMP, SCH, 2>;
        else return nullptr;
    }
};
struct GbmAnti {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return FN == FN_MIN ? nullptr : level_anti_kernel<R, FN == FN_MIN ? FN_FINAL : FN, MT, SAMP, SCH>; }
};
#elif MLMCB_TU_FAMILY == 1
struct MertonUniform {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return level_kernel<R, 1, FN, MT, SAMP, SCH>; }
};
#elif MLMCB_TU_FAMILY == 2
struct MertonFlat {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return mjd_flat_kernel<R, FN, MT, SAMP, SCH>; }
};
struct MertonBatch {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return mjd_batch_kernel<R, FN, MT, SAMP, SCH>; }
};
#endif
}  // namespace

#if MLMCB_TU_FAMILY == 0
template <> LevelKernel pick_gbm_s<S>(int fn, int mt, int flags) { return pick_family<Gbm, S>(fn, mt, flags); }
template <> LevelKernel pick_gbm_packed_s<S>(int fn, int flags, int nsteps) {
    // nsteps 1 (l = 0) and 2 (M = 2, l = 1) share one kernel; 4 steps come as M = 4, l = 1 (mt 4) or M = 2, l = 2 (mt 2)
    if (nsteps == 4) return pick_family<GbmPacked<4>, S>(fn, 4, flags);
    if (nsteps == -4) return pick_family<GbmPacked<-4>, S>(fn, 2, flags);
    return pick_family<GbmPacked<0>, S>(fn, 0, flags);
}
template <> LevelKernel pick_gbm_long_s<S>(int fn, int mt) { return pick_family<GbmLong, S>(fn, mt, 0); }
template <> LevelKernel pick_gbm_anti_s<S>(int fn, int mt, int flags) { return pick_family<GbmAnti, S>(fn, mt, flags); }
#elif MLMCB_TU_FAMILY == 1
template <> LevelKernel pick_merton_uniform_s<S>(int fn, int mt, int flags) { return pick_family<MertonUniform, S>(fn, mt, flags); }
#elif MLMCB_TU_FAMILY == 2
template <> LevelKernel pick_merton_flat_s<S>(int fn, int mt, int flags, bool batch) {
    return batch ? pick_family<MertonBatch, S>(fn, mt, flags) : pick_family<MertonFlat, S>(fn, mt, flags);
}
#elif MLMCB_TU_FAMILY == 5
template <> LevelKernel pick_merton_factor_s<S>(int fn, int mt) {
    if (mt == 2) return fn == FN_AVG ? mjd_factor_kernel<FN_AVG, 2, S> : fn == FN_MIN ? mjd_factor_kernel<FN_MIN, 2, S> : mjd_factor_kernel<FN_FINAL, 2, S>;
    return fn == FN_AVG ? mjd_factor_kernel<FN_AVG, 4, S> : fn == FN_MIN ? mjd_factor_kernel<FN_MIN, 4, S> : mjd_factor_kernel<FN_FINAL, 4, S>;
}
#elif MLMCB_TU_FAMILY == 3
template <> SweepKernel pick_gbm_sweep_s<S>(int fn, int mt, int flags) { return pick_sweep<0, S>(fn, mt, flags); }
#else
template <> SweepKernel pick_merton_sweep_s<S>(int fn, int mt, int flags) { return pick_sweep<1, S>(fn, mt, flags); }
#endif
}  // namespace mlmcb
