// iterate_launch_decl.cuh -- declarations of the per-translation-unit launchers (see iterate_launch.cuh)
#pragma once
#include "kernels.cuh"
namespace admmb {
struct IterLaunchCtx {
    cudaStream_t stream;
    int num_sms;
    int N, nb;
    bool par_batched, has_c, has_q, fast_pattern, decoupled;
    bool two_per_thread;          // use the two-problems-per-thread kernel (iterate2.cuh) when eligible
};
void launch_iterate_smem(const IterLaunchCtx &c, const IterParams &P, bool adapt);
void launch_iterate_gshared(const IterLaunchCtx &c, const IterParams &P, bool adapt);
void launch_iterate_pp(const IterLaunchCtx &c, const IterParams &P, bool adapt);
// per-problem factors staged through shared memory by TMA; false when the configuration is not eligible
bool launch_iterate_pptma(const IterLaunchCtx &c, const IterParams &P, bool adapt);
}  // namespace admmb
