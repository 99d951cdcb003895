// iter_pp.cu -- one third of the k_admm_iterate template variants (see iterate_launch.cuh)
#include "iterate_launch.cuh"
namespace admmb {
void launch_iterate_pp(const IterLaunchCtx &c, const IterParams &P, bool adapt) { launch_iterate_tu<false, false>(c, P, adapt); }
}  // namespace admmb
