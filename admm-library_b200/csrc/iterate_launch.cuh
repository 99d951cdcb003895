// iterate_launch.cuh -- launch logic of the persistent ADMM kernel.  The kernel has 3 x 8 x 3 x 2 template
// variants; they are instantiated in three translation units (iter_smem.cu, iter_gshared.cu, iter_pp.cu),
// compiled in parallel, so that a rebuild takes about a minute instead of four.
#pragma once
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "kernels.cuh"
#include "iterate_launch_decl.cuh"

namespace admmb {

// smallest CTA that still puts every active problem on the machine in one wave; otherwise the CTA
// size with the largest resident capacity.  Returns the resident capacity in *cap_out.
template <class K>
static int pick_block(K kern, size_t smem, int num_sms, int n_active, long *cap_out)
{
    int bestT = 128;
    long best_cap = -1;
    for (int T = 32; T <= 256; T += 32) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem) != cudaSuccess || occ <= 0) continue;
        long cap = (long)occ * num_sms * T;
        if ((long)n_active <= cap) { *cap_out = cap; return T; }
        if (cap > best_cap) { best_cap = cap; bestT = T; }
    }
    *cap_out = best_cap;
    return bestT;
}

template <class K1, class K2>
static void launch_iterate_kernel(const IterLaunchCtx &c, K1 kern, K2 kern_lowocc, const IterParams &P, size_t smem)
{
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern_lowocc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the uncapped-register build when it still holds the whole active set in one wave
    long cap_lo = 0, cap_hi = 0;
    const int T_lo = pick_block(kern_lowocc, smem, c.num_sms, P.n_active, &cap_lo);
    if ((long)P.n_active <= cap_lo) {
        kern_lowocc<<<(P.n_active + T_lo - 1) / T_lo, T_lo, smem, c.stream>>>(P);
    } else {
        const int T = pick_block(kern, smem, c.num_sms, P.n_active, &cap_hi);
        kern<<<(P.n_active + T - 1) / T, T, smem, c.stream>>>(P);
    }
    CK(cudaGetLastError());
}

template <bool FSH, bool FSMEM>
static void launch_iterate_tu(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    size_t smem = 16 + ((FSH && FSMEM) ? sizeof(double) * (c.decoupled ? FD : FS) * c.N : 0) +
                  (c.par_batched ? 0 : sizeof(double) * 8 * c.nb) + sizeof(int) * ((c.nb + 3) / 4) * 4;
    smem = round_up(smem, 16);
#define DISPATCH(C, Q, A)                                                                              \
    do {                                                                                              \
        if (c.fast_pattern && c.decoupled)                                                            \
            launch_iterate_kernel(c, k_admm_iterate<FSH, FSMEM, C, Q, A, 2, false>,                   \
                                  k_admm_iterate<FSH, FSMEM, C, Q, A, 2, true>, P, smem);             \
        else if (c.fast_pattern)                                                                      \
            launch_iterate_kernel(c, k_admm_iterate<FSH, FSMEM, C, Q, A, 1, false>,                   \
                                  k_admm_iterate<FSH, FSMEM, C, Q, A, 1, true>, P, smem);             \
        else                                                                                          \
            launch_iterate_kernel(c, k_admm_iterate<FSH, FSMEM, C, Q, A, 0, false>,                   \
                                  k_admm_iterate<FSH, FSMEM, C, Q, A, 0, false>, P, smem);            \
    } while (0)
    if (c.has_c) {
        if (c.has_q) { if (adapt) DISPATCH(true, true, true); else DISPATCH(true, true, false); }
        else { if (adapt) DISPATCH(true, false, true); else DISPATCH(true, false, false); }
    } else {
        if (c.has_q) { if (adapt) DISPATCH(false, true, true); else DISPATCH(false, true, false); }
        else { if (adapt) DISPATCH(false, false, true); else DISPATCH(false, false, false); }
    }
#undef DISPATCH
}

}  // namespace admmb
