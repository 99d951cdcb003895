// iter_smem.cu -- one third of the k_admm_iterate template variants (see iterate_launch.cuh)
#define ADMMB_WITH_ITERATE2
#include "iterate_launch.cuh"
namespace admmb {
void launch_iterate_smem(const IterLaunchCtx &c, const IterParams &P, bool adapt) { launch_iterate_tu<true, true>(c, P, adapt); }
}  // namespace admmb
