// iterate_launch.cuh -- launch logic of the persistent ADMM kernel.  The kernel has 3 x 8 x 3 x 2 template
// variants; they are instantiated in three translation units (iter_smem.cu, iter_gshared.cu, iter_pp.cu),
// compiled in parallel, so that a rebuild takes about a minute instead of four.
#pragma once
#include <cstdlib>
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "kernels.cuh"
#include "iterate_launch_decl.cuh"
#ifdef ADMMB_WITH_ITERATE2
#include "iterate2.cuh"
#endif

namespace admmb {

// Resident capacity of one kernel variant for CTA sizes 32..256, queried once and cached: the launch
// loop runs hundreds of times per solve with the GPU idle while the host decides.
struct OccTable {
    const void *kern = nullptr;
    size_t smem = 0;
    int num_sms = 0;
    int device = -1;              // cudaFuncSetAttribute and the occupancy query are per device
    long cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

template <class K>
static const OccTable &occ_table(K kern, size_t smem, int num_sms, int device)
{
    static thread_local OccTable cache[32];
    static thread_local int used = 0;
    // the dynamic-smem limit is per-function state: it must cover the largest request made so far
    size_t attr = 0;
    const OccTable *hit = nullptr;
    for (int i = 0; i < used; ++i)
        if (cache[i].kern == (const void *)kern && cache[i].device == device) {
            attr = attr > cache[i].smem ? attr : cache[i].smem;
            if (cache[i].smem == smem && cache[i].num_sms == num_sms) hit = &cache[i];
        }
    if (hit) return *hit;
    if (smem > attr) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (used == 32) used = 0;   // tiny cache: start over (re-queries, never wrong: attr is re-derived from live entries)
    OccTable &t = cache[used++];
    t.kern = (const void *)kern; t.smem = smem; t.num_sms = num_sms; t.device = device;
    for (int i = 0; i < 8; ++i) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * (i + 1), smem) != cudaSuccess) occ = 0;
        t.cap[i] = (long)occ * num_sms * 32 * (i + 1);
    }
    return t;
}

// smallest CTA that still puts every active problem on the machine in one wave; otherwise the CTA
// size with the largest resident capacity.
static int pick_block(const OccTable &t, int n_active, bool *one_wave)
{
    int best = 3;
    static const int min_i = getenv("ADMMB_MIN_BLOCK") ? atoi(getenv("ADMMB_MIN_BLOCK")) / 32 - 1 : 0;   // tuning knob
    for (int i = min_i; i < 8; ++i) {
        if (t.cap[i] >= (long)n_active) { *one_wave = true; return 32 * (i + 1); }
        if (t.cap[i] > t.cap[best]) best = i;
    }
    *one_wave = false;
    return 32 * (best + 1);
}

template <class K1, class K2>
static void launch_iterate_kernel(const IterLaunchCtx &c, K1 kern, K2 kern_lowocc, const IterParams &P, size_t smem)
{
    // the uncapped-register build when it still holds the whole active set in one wave
    bool one_wave = false;
    const int T_lo = pick_block(occ_table(kern_lowocc, smem, c.num_sms, c.device), P.n_active, &one_wave);
    if (c.kernel == KV_THREAD) one_wave = false;            // pinned: the register-capped build
    if (c.kernel == KV_THREAD_WIDE) one_wave = true;        // pinned: the uncapped build, however many waves
    if (one_wave) {
        kern_lowocc<<<(P.n_active + T_lo - 1) / T_lo, T_lo, smem, c.stream>>>(P);
    } else {
        const int T = pick_block(occ_table(kern, smem, c.num_sms, c.device), P.n_active, &one_wave);
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
#ifdef ADMMB_WITH_ITERATE2
    // Two problems per thread (iterate2.cuh): +8 % at full width (half the LDS traffic per problem), but a lone
    // warp gains nothing from the second stream (57 us vs 32 us per iteration for twice the work), so it is used
    // only when the working set is too wide for the uncapped one-problem build to hold it in one wave.
    if (FSH && FSMEM && (c.kernel == KV_AUTO || c.kernel == KV_THREAD2) && c.two_per_thread && c.fast_pattern && c.decoupled &&
        !c.has_c && !c.has_q && !c.par_batched) {
        bool lo_fits = false, one_wave = false;
        if (adapt) pick_block(occ_table(k_admm_iterate<FSH, FSMEM, false, false, true, 2, true>, smem, c.num_sms, c.device), P.n_active, &lo_fits);
        else pick_block(occ_table(k_admm_iterate<FSH, FSMEM, false, false, false, 2, true>, smem, c.num_sms, c.device), P.n_active, &lo_fits);
        if (c.kernel == KV_THREAD2) lo_fits = false;        // pinned
        if (!lo_fits) {
            const int half = (P.n_active + 1) / 2;
            if (adapt) {
                const int T = pick_block(occ_table(k_admm_iterate2<true>, smem, c.num_sms, c.device), half, &one_wave);
                k_admm_iterate2<true><<<(P.n_active + 2 * T - 1) / (2 * T), T, smem, c.stream>>>(P);
            } else {
                const int T = pick_block(occ_table(k_admm_iterate2<false>, smem, c.num_sms, c.device), half, &one_wave);
                k_admm_iterate2<false><<<(P.n_active + 2 * T - 1) / (2 * T), T, smem, c.stream>>>(P);
            }
            CK(cudaGetLastError());
            return;
        }
    }
#endif
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
