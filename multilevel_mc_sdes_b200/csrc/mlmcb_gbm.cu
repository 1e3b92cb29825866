// libmlmcb - GBM level kernels (level_kernel<R, 0, ...>; see mlmcb_kernels.cuh).
#define MLMCB_TEMPLATES_ONLY
#include "mlmcb_pick.cuh"

namespace mlmcb {
namespace {
struct Gbm {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return level_kernel<R, 0, FN, MT, SAMP, SCH>; }
};
}  // namespace
LevelKernel pick_gbm(int fn, int mt, int flags) { return pick_family<Gbm>(fn, mt, flags); }
}  // namespace mlmcb
