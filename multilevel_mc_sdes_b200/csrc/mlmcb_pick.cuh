// libmlmcb - selecting one instantiation of a level-kernel family from run-time codes.
// A family x sampler pair is one translation unit: mlmcb_family.cu compiled with
// -DMLMCB_TU_FAMILY={0 GBM, 1 Merton block-uniform, 2 Merton event-driven, 3 GBM sweep, 4 Merton sweep, 5 Merton factor walk}
// -DMLMCB_TU_SAMP={0 f32bm, 1 f64bm} (build.py), so nvcc compiles the twelve of them side by side; replay, American and probe kernels stay in
// mlmcb_api.cu.
#pragma once
#include "../../include/mlmcb.h"
#include "mlmcb_kernels.cuh"

namespace mlmcb {

typedef void (*LevelKernel)(const LevelConsts, double*, unsigned*, double*, double*, double*);
typedef void (*SweepKernel)(const SweepConsts, double*, unsigned*, double*);

// Fam::get<R, FN, MT, SAMP, SCH>() returns the kernel
template <class Fam, typename R, int FN, int SAMP, int SCH>
LevelKernel pick_mt(int mt) {
    switch (mt) {
    case 2: return Fam::template get<R, FN, 2, SAMP, SCH>();
    case 4: return Fam::template get<R, FN, 4, SAMP, SCH>();
    default: return Fam::template get<R, FN, 0, SAMP, SCH>();
    }
}
template <class Fam, typename R, int SAMP, int SCH = SCH_EM>
LevelKernel pick_fn(int fn, int mt) {
    switch (fn) {
    case FN_AVG: return pick_mt<Fam, R, FN_AVG, SAMP, SCH>(mt);
    case FN_MIN: return pick_mt<Fam, R, FN_MIN, SAMP, SCH>(mt);
    default: return pick_mt<Fam, R, FN_FINAL, SAMP, SCH>(mt);
    }
}
// fn: FN_*; mt: 2, 4 or 0 (generic M); flags: MLMCB_FP32_PATHS, MLMCB_MILSTEIN.  Per sampler: fp64 paths with
// either scheme, fp32 paths with Euler-Maruyama (check_flags refuses Milstein in fp32) - 27 kernels.
template <class Fam, int SAMP>
LevelKernel pick_family(int fn, int mt, int flags) {
    if (flags & MLMCB_MILSTEIN) return pick_fn<Fam, double, SAMP, SCH_MILSTEIN>(fn, mt);
    if (flags & MLMCB_FP32_PATHS) return pick_fn<Fam, float, SAMP>(fn, mt);
    return pick_fn<Fam, double, SAMP>(fn, mt);
}

// the sweep kernels: fp64 paths, plain estimator, either scheme (anything else takes the per-level launches)
template <int MODEL, int FN, int SAMP, int SCH>
SweepKernel pick_sweep_mt(int mt) {
    switch (mt) {
    case 2: return sweep_kernel<double, MODEL, FN, 2, SAMP, SCH>;
    case 4: return sweep_kernel<double, MODEL, FN, 4, SAMP, SCH>;
    default: return sweep_kernel<double, MODEL, FN, 0, SAMP, SCH>;
    }
}
template <int MODEL, int SAMP>
SweepKernel pick_sweep(int fn, int mt, int flags) {
    if (flags & MLMCB_MILSTEIN)
        return fn == FN_AVG ? pick_sweep_mt<MODEL, FN_AVG, SAMP, SCH_MILSTEIN>(mt)
             : fn == FN_MIN ? pick_sweep_mt<MODEL, FN_MIN, SAMP, SCH_MILSTEIN>(mt) : pick_sweep_mt<MODEL, FN_FINAL, SAMP, SCH_MILSTEIN>(mt);
    return fn == FN_AVG ? pick_sweep_mt<MODEL, FN_AVG, SAMP, SCH_EM>(mt)
         : fn == FN_MIN ? pick_sweep_mt<MODEL, FN_MIN, SAMP, SCH_EM>(mt) : pick_sweep_mt<MODEL, FN_FINAL, SAMP, SCH_EM>(mt);
}

// defined (as explicit specialisations) by the translation unit of that family and sampler
template <int SAMP> LevelKernel pick_gbm_s(int fn, int mt, int flags);
template <int SAMP> LevelKernel pick_gbm_packed_s(int fn, int flags, int nsteps);   // paths of 1, 2 or 4 fine steps
template <int SAMP> LevelKernel pick_gbm_anti_s(int fn, int mt, int flags);         // antithetic triple (Euro/Asian)
template <int SAMP> LevelKernel pick_gbm_long_s(int fn, int mt);                    // long paths, fp64 sampler (nullptr otherwise)
template <int SAMP> LevelKernel pick_merton_uniform_s(int fn, int mt, int flags);
template <int SAMP> LevelKernel pick_merton_flat_s(int fn, int mt, int flags, bool batch);   // batch: the batch walk
template <int SAMP> LevelKernel pick_merton_factor_s(int fn, int mt);                        // MLMCB_TU_FAMILY 5: factor walk (EM, fp64 paths, mt 2 | 4)
template <int SAMP> SweepKernel pick_gbm_sweep_s(int fn, int mt, int flags);         // MLMCB_TU_FAMILY 3
template <int SAMP> SweepKernel pick_merton_sweep_s(int fn, int mt, int flags);      // MLMCB_TU_FAMILY 4

}  // namespace mlmcb
