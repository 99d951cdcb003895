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
    if (!enabled || !c.fast_pattern) return false;
    const bool gen = !c.decoupled;    // generic 156-double records (e.g. the SCP linearisations of scp.cuh)
    size_t smem = 16 + (c.par_batched ? 0 : sizeof(double) * 8 * c.nb) + sizeof(int) * ((c.nb + 3) / 4) * 4 + PPT_SLOTS * 8 +
                  128 + (size_t)PPT_SLOTS * ppt_slot_bytes(c.has_c, gen);
    smem = round_up(smem, 128);
    if (smem > 227 * 1024) return false;     // long horizons with a shared parameter table: the strided-load kernel
    PpTmaMaps maps;
    if (gen) {
        const size_t rows = (size_t)FS * c.N;
        maps.mB = tmap_rows_f64(P.fac, rows, P.ld, F_A);
        maps.mF0 = tmap_rows_f64(P.fac, rows, P.ld, 18);
        maps.mF1 = tmap_rows_f64(P.fac, rows, P.ld, 66);
        maps.m6 = tmap_rows_f64(P.fac, rows, P.ld, 6);
    } else {
        const size_t rows = (size_t)FD * c.N;
        maps.mB = tmap_rows_f64(P.fac_dec, rows, P.ld, 46);
        maps.mF0 = tmap_rows_f64(P.fac_dec, rows, P.ld, 10);
        maps.mF1 = tmap_rows_f64(P.fac_dec, rows, P.ld, 30);
        maps.m6 = tmap_rows_f64(P.fac_dec, rows, P.ld, 6);
    }
    // One warp per CTA with an eight-slot ring once every warp can have an SM (almost) to itself: a lone warp's
    // iteration is bound by the DRAM latency of its stage records, which seven requests in flight cover and one does not.
    static const int deep_width = getenv("ADMMB_PPT_DEEP") ? atoi(getenv("ADMMB_PPT_DEEP")) : 32 * c.num_sms;   // tuning knob
    // (measured, config 4, N = 50: 1,024 problems 59 -> 52 us per iteration, 4,096: 82 -> 71 us; two such CTAs on an SM are slower
    // than the four-warp form: 8,192 problems 92 -> 140 us)
    const bool deep = P.n_active <= deep_width;
    const int T = deep ? 32 : 128;
    const unsigned grid = (unsigned)((P.n_active + T - 1) / T);
#define PPT_LAUNCH1(C, Q, A, W, G)                                                                                   \
    do {                                                                                                             \
        CK(cudaFuncSetAttribute(k_admm_iterate_pptma<C, Q, A, W, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_admm_iterate_pptma<C, Q, A, W, G><<<grid, T, smem, c.stream>>>(P, maps);                                   \
    } while (0)
#define PPT_LAUNCH(C, Q, A)                                                                                          \
    do {                                                                                                             \
        if (gen) { if (deep) PPT_LAUNCH1(C, Q, A, 1, true); else PPT_LAUNCH1(C, Q, A, 4, true); }                    \
        else { if (deep) PPT_LAUNCH1(C, Q, A, 1, false); else PPT_LAUNCH1(C, Q, A, 4, false); }                      \
    } while (0)
    if (c.has_c) {
        if (c.has_q) { if (adapt) PPT_LAUNCH(true, true, true); else PPT_LAUNCH(true, true, false); }
        else { if (adapt) PPT_LAUNCH(true, false, true); else PPT_LAUNCH(true, false, false); }
    } else {
        if (c.has_q) { if (adapt) PPT_LAUNCH(false, true, true); else PPT_LAUNCH(false, true, false); }
        else { if (adapt) PPT_LAUNCH(false, false, true); else PPT_LAUNCH(false, false, false); }
    }
#undef PPT_LAUNCH
#undef PPT_LAUNCH1
    CK(cudaGetLastError());
    return true;
}

}  // namespace admmb
