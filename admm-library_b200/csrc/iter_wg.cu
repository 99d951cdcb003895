// iter_wg.cu -- launcher of the warp-group kernel (iterate_wg.cuh): tile width from the shared-memory budget,
// grid = one CTA per resident tile (persistent over tiles when there are more tiles than SMs).
#include <cstdlib>
#include "host_util.cuh"
#include "tma_host.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_wg.cuh"
#include "iterate_launch_decl.cuh"

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

// widest tile (problems per CTA, <= 32) whose z, u, d, g rows fit next to the factor; 0: not eligible
int iterate_wg_tile_width(const IterLaunchCtx &c)
{
    if (!c.fast_pattern || !c.decoupled || c.has_c || c.has_q || c.par_batched || c.rows_zu <= 0) return 0;
    static const int tw_cap = getenv("ADMMB_WG_TW") ? atoi(getenv("ADMMB_WG_TW")) : 32;      // developer knob
    for (int tw = tw_cap < 32 ? tw_cap : 32; tw >= WG_MIN_TW; --tw)
        if (wg_layout(c.N, c.rows_zu, tw, c.time_invariant ? D_AIN : FD).total <= WG_SMEM_MAX) return tw;
    return 0;
}

// shared factor only (the caller checks)
bool launch_iterate_wg(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    const int tw = iterate_wg_tile_width(c);
    if (tw == 0) return false;
    const size_t smem = wg_layout(c.N, c.rows_zu, tw, c.time_invariant ? D_AIN : FD).total;
    const int ntiles = (P.n_active + tw - 1) / tw;
    const int grid = ntiles < c.num_sms ? ntiles : c.num_sms;
    static const WgPpMaps no_maps = {};
#define WG_LAUNCH(A, T)                                                                   \
    do {                                                                                  \
        wg_set_attr(k_admm_iterate_wg<A, T>, smem, c.device);                             \
        k_admm_iterate_wg<A, T><<<grid, WG_WARPS * 32, smem, c.stream>>>(P, tw, no_maps); \
    } while (0)
    if (adapt) { if (c.time_invariant) WG_LAUNCH(true, true); else WG_LAUNCH(true, false); }
    else { if (c.time_invariant) WG_LAUNCH(false, true); else WG_LAUNCH(false, false); }
#undef WG_LAUNCH
    CK(cudaGetLastError());
    return true;
}

// ---- per-problem models (config 4): the same kernel with the stage records streamed through a TMA ring (iterate_wg.cuh, PP)
// widest even tile whose z, u, d, g rows fit next to the ring; 0: not eligible.  A rho change must not need a new factor
// (no quadratic cost), the affine term / linear cost / per-problem parameters are not handled (as in the shared-factor form).
int iterate_wgpp_tile_width(const IterLaunchCtx &c, bool refactors)
{
    if (!c.fast_pattern || !c.decoupled || c.has_c || c.has_q || c.par_batched || c.rows_zu <= 0 || refactors) return 0;
    static const int tw_cap = getenv("ADMMB_WG_TW") ? atoi(getenv("ADMMB_WG_TW")) : 32;
    for (int tw = (tw_cap < 32 ? tw_cap : 32) & ~1; tw >= WG_MIN_TW; tw -= 2)
        if (wg_layout(c.N, c.rows_zu, tw, 0, true).total <= WG_SMEM_MAX) return tw;
    return 0;
}

bool launch_iterate_wgpp(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    const int tw = iterate_wgpp_tile_width(c, adapt && P.has_P);
    if (tw == 0) return false;
    const size_t smem = wg_layout(c.N, c.rows_zu, tw, 0, true).total;
    const int ntiles = (P.n_active + tw - 1) / tw;
    const int grid = ntiles < c.num_sms ? ntiles : c.num_sms;
    const size_t rows = (size_t)FD * c.N;
    WgPpMaps maps;
    maps.mB = tmap_box_f64(P.fac_dec, rows, P.ld, WG_PP_ROWS, (uint32_t)tw);
    maps.mF0 = tmap_box_f64(P.fac_dec, rows, P.ld, 10, (uint32_t)tw);
    maps.mF1 = tmap_box_f64(P.fac_dec, rows, P.ld, 30, (uint32_t)tw);
    if (adapt) {
        wg_set_attr(k_admm_iterate_wg<true, false, true>, smem, c.device);
        k_admm_iterate_wg<true, false, true><<<grid, WG_WARPS * 32, smem, c.stream>>>(P, tw, maps);
    } else {
        wg_set_attr(k_admm_iterate_wg<false, false, true>, smem, c.device);
        k_admm_iterate_wg<false, false, true><<<grid, WG_WARPS * 32, smem, c.stream>>>(P, tw, maps);
    }
    CK(cudaGetLastError());
    return true;
}

}  // namespace admmb
