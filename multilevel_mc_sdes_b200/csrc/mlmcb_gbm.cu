// libmlmcb - GBM level kernels (level_kernel<R, 0, ...> and gbm_packed_kernel; see mlmcb_kernels.cuh).
#define MLMCB_TEMPLATES_ONLY
#include "mlmcb_pick.cuh"

namespace mlmcb {
namespace {
struct Gbm {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return level_kernel<R, 0, FN, MT, SAMP, SCH>; }
};
struct GbmPacked {      // one or two fine steps: the coarseness M plays no role
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return gbm_packed_kernel<R, FN, SAMP, SCH>; }
};
}  // namespace
LevelKernel pick_gbm(int fn, int mt, int flags) { return pick_family<Gbm>(fn, mt, flags); }
LevelKernel pick_gbm_packed(int fn, int flags) { return pick_family<GbmPacked>(fn, 0, flags); }
}  // namespace mlmcb
