// iter_wg_host.cuh -- host helpers shared by the launchers of the warp-group kernel (iter_wg.cu, iter_wgpp.cu)
#pragma once
#include "host_util.cuh"

namespace admmb {

namespace {
constexpr size_t WG_SMEM_MAX = 227 * 1024;      // opt-in dynamic shared memory per CTA on sm_100

struct WgAttr { const void *kern; int device; size_t smem; };

template <class K>
void wg_set_attr(K kern, size_t smem, int device)
{
    static thread_local WgAttr done[8];
    static thread_local int used = 0;
    for (int i = 0; i < used; ++i)
        if (done[i].kern == (const void *)kern && done[i].device == device && done[i].smem >= smem) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (used == 8) used = 0;
    done[used++] = WgAttr{(const void *)kern, device, smem};
}
}  // namespace

}  // namespace admmb
