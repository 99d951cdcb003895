// iter_pptma.cu -- launcher of the TMA-staged per-problem-factor kernel (iterate_pptma.cuh)
#include <cstdlib>
#include "host_util.cuh"
#include "tma_host.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_pptma.cuh"
#include "iterate_launch_decl.cuh"

namespace admmb {


bool launch_iterate_pptma(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    static const bool enabled = getenv("ADMMB_PPTMA") ? atoi(getenv("ADMMB_PPTMA")) != 0 : true;
    if (!enabled || !c.fast_pattern || !c.decoupled) return false;
    const size_t rows = (size_t)FD * c.N;
    PpTmaMaps maps;
    maps.m46 = tmap_rows_f64(P.fac_dec, rows, P.ld, 46);
    maps.m10 = tmap_rows_f64(P.fac_dec, rows, P.ld, 10);
    maps.m30 = tmap_rows_f64(P.fac_dec, rows, P.ld, 30);
    maps.m6 = tmap_rows_f64(P.fac_dec, rows, P.ld, 6);
    size_t smem = 16 + (c.par_batched ? 0 : sizeof(double) * 8 * c.nb) + sizeof(int) * ((c.nb + 3) / 4) * 4 + PPT_SLOTS * 8 +
                  128 + (size_t)PPT_SLOTS * ppt_slot_bytes(c.has_c);
    smem = round_up(smem, 128);
    // One warp per CTA with an eight-slot ring once every warp can have an SM (almost) to itself: a lone warp's
    // iteration is bound by the DRAM latency of its stage records, which seven requests in flight cover and one does not.
    static const int deep_width = getenv("ADMMB_PPT_DEEP") ? atoi(getenv("ADMMB_PPT_DEEP")) : 32 * c.num_sms;   // tuning knob
    // (measured, config 4, N = 50: 1,024 problems 59 -> 52 us per iteration, 4,096: 82 -> 71 us; two such CTAs on an SM are slower
    // than the four-warp form: 8,192 problems 92 -> 140 us)
    const bool deep = P.n_active <= deep_width;
    const int T = deep ? 32 : 128;
    const unsigned grid = (unsigned)((P.n_active + T - 1) / T);
#define PPT_LAUNCH(C, Q, A)                                                                                          \
    do {                                                                                                             \
        if (deep) {                                                                                                  \
            CK(cudaFuncSetAttribute(k_admm_iterate_pptma<C, Q, A, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_admm_iterate_pptma<C, Q, A, 1><<<grid, T, smem, c.stream>>>(P, maps);                                  \
        } else {                                                                                                     \
            CK(cudaFuncSetAttribute(k_admm_iterate_pptma<C, Q, A, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_admm_iterate_pptma<C, Q, A, 4><<<grid, T, smem, c.stream>>>(P, maps);                                  \
        }                                                                                                            \
    } while (0)
    if (c.has_c) {
        if (c.has_q) { if (adapt) PPT_LAUNCH(true, true, true); else PPT_LAUNCH(true, true, false); }
        else { if (adapt) PPT_LAUNCH(true, false, true); else PPT_LAUNCH(true, false, false); }
    } else {
        if (c.has_q) { if (adapt) PPT_LAUNCH(false, true, true); else PPT_LAUNCH(false, true, false); }
        else { if (adapt) PPT_LAUNCH(false, false, true); else PPT_LAUNCH(false, false, false); }
    }
#undef PPT_LAUNCH
    CK(cudaGetLastError());
    return true;
}

}  // namespace admmb
