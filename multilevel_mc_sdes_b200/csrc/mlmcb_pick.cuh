// libmlmcb - selecting one instantiation of a level-kernel family from run-time codes.
// The families live in separate translation units (mlmcb_gbm.cu, mlmcb_merton_uniform.cu,
// mlmcb_merton_flat.cu; antithetic, replay, American and probe kernels stay in mlmcb_api.cu) so that
// nvcc compiles them side by side.
#pragma once
#include "../../include/mlmcb.h"
#include "mlmcb_kernels.cuh"

namespace mlmcb {

typedef void (*LevelKernel)(const LevelConsts, double*, unsigned*, double*, double*, double*);

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
// fn: FN_*; mt: 2, 4 or 0 (generic M); flags: sampler, MLMCB_FP32_PATHS, MLMCB_MILSTEIN - 36 kernels per family
template <class Fam>
LevelKernel pick_family(int fn, int mt, int flags) {
    const int samp = flags & MLMCB_SAMPLER_MASK;
    if (flags & MLMCB_MILSTEIN) return pick_fn<Fam, double, SAMP_F32BM, SCH_MILSTEIN>(fn, mt);
    if (flags & MLMCB_FP32_PATHS) return pick_fn<Fam, float, SAMP_F32BM>(fn, mt);      // default sampler only (check_flags)
    return samp == MLMCB_SAMPLER_F64BM ? pick_fn<Fam, double, SAMP_F64BM>(fn, mt) : pick_fn<Fam, double, SAMP_F32BM>(fn, mt);
}

LevelKernel pick_gbm(int fn, int mt, int flags);              // mlmcb_gbm.cu
LevelKernel pick_gbm_packed(int fn, int flags);               // mlmcb_gbm.cu: paths of one or two fine steps
LevelKernel pick_merton_uniform(int fn, int mt, int flags);   // mlmcb_merton_uniform.cu
LevelKernel pick_merton_flat(int fn, int mt, int flags);      // mlmcb_merton_flat.cu

}  // namespace mlmcb
