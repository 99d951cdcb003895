// iter_wg.cu -- launcher of the warp-group kernel (iterate_wg.cuh): tile width from the shared-memory budget,
// grid = one CTA per resident tile (persistent over tiles when there are more tiles than SMs).
#include <cstdlib>
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_wg.cuh"
#include "iterate_launch_decl.cuh"
#include "iter_wg_host.cuh"

namespace admmb {


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

}  // namespace admmb
