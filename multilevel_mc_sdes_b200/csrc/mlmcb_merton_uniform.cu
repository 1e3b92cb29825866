// libmlmcb - Merton level kernels, block-uniform walk (level_kernel<R, 1, ...>; see mlmcb_kernels.cuh).
#define MLMCB_TEMPLATES_ONLY
#include "mlmcb_pick.cuh"

namespace mlmcb {
namespace {
struct MertonUniform {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return level_kernel<R, 1, FN, MT, SAMP, SCH>; }
};
}  // namespace
LevelKernel pick_merton_uniform(int fn, int mt, int flags) { return pick_family<MertonUniform>(fn, mt, flags); }
}  // namespace mlmcb
