// libmlmcb - Merton level kernels, event-driven walk (mjd_flat_kernel; see mlmcb_kernels.cuh).
#define MLMCB_TEMPLATES_ONLY
#include "mlmcb_pick.cuh"

namespace mlmcb {
namespace {
struct MertonFlat {
    template <typename R, int FN, int MT, int SAMP, int SCH>
    static LevelKernel get() { return mjd_flat_kernel<R, FN, MT, SAMP, SCH>; }
};
}  // namespace
LevelKernel pick_merton_flat(int fn, int mt, int flags) { return pick_family<MertonFlat>(fn, mt, flags); }
}  // namespace mlmcb
