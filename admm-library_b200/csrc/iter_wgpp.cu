// iter_wgpp.cu -- launcher of the warp-group kernel for per-problem models (iterate_wg.cuh, PP): the stage records of a
// tile are streamed through a TMA ring; one instantiation per tile width (8, 16, 24, 32 problems).
#include <cstdlib>
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_wg.cuh"
#include "iterate_launch_decl.cuh"
#include "iter_wg_host.cuh"

namespace admmb {

// ---- per-problem models (config 4): the same kernel with the stage records streamed through a TMA ring (iterate_wg.cuh, PP)
// widest tile of 8, 16, 24 or 32 problems whose z, u, d, g rows fit next to the ring; 0: not eligible.  Multiples of 8: one
// instantiation per width (the first form, TMA boxes, needed them for the 128-byte alignment of every box; a bulk copy needs
// 16 bytes, which any even width gives).  A rho change must not need a new factor
// (no quadratic cost), the affine term / linear cost / per-problem parameters are not handled (as in the shared-factor form).
int iterate_wgpp_tile_width(const IterLaunchCtx &c, bool refactors)
{
    if (!c.fast_pattern || !c.decoupled || c.has_c || c.has_q || c.par_batched || c.rows_zu <= 0 || refactors) return 0;
    static const int tw_cap = getenv("ADMMB_WG_TW") ? atoi(getenv("ADMMB_WG_TW")) : 32;
    static_assert(WG_MIN_TW % 8 == 0, "instantiated widths");
    for (int tw = (tw_cap < 32 ? tw_cap : 32) & ~7; tw >= WG_MIN_TW; tw -= 8)
        if (wg_layout(c.N, c.rows_zu, tw, 0, true).total <= WG_SMEM_MAX) return tw;
    return 0;
}

size_t wgpp_block_doubles(const IterLaunchCtx &c, int n_active, int tw)
{
    const size_t ntiles = ((size_t)n_active + tw - 1) / tw;
    return ntiles * (size_t)c.N * WG_PP_BLK_ROWS * tw;
}

bool launch_iterate_wgpp(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    const int tw = iterate_wgpp_tile_width(c, adapt && P.has_P);
    if (tw == 0) return false;
    const size_t smem = wg_layout(c.N, c.rows_zu, tw, 0, true).total;
    const int ntiles = (P.n_active + tw - 1) / tw;
    const int grid = ntiles < c.num_sms ? ntiles : c.num_sms;
    // the records of this launch's working set, tile by tile (a repack between launches moves columns, so the copy is made
    // per launch: one pass over the records, a few per cent of a launch of 50+ iterations)
    if (!c.wgpp_blk) return false;
    k_wgpp_block<<<dim3((unsigned)ntiles, (unsigned)c.N), 256, 0, c.stream>>>(P.fac_dec, P.ld, c.N, P.n_active, tw, c.wgpp_blk);
    WgPpMaps maps = {c.wgpp_blk};
#define WGPP_LAUNCH(A, W)                                                                              \
    do {                                                                                               \
        wg_set_attr(k_admm_iterate_wg<A, false, true, W>, smem, c.device);                             \
        k_admm_iterate_wg<A, false, true, W><<<grid, WG_WARPS * 32, smem, c.stream>>>(P, tw, maps);    \
    } while (0)
#define WGPP_WIDTH(W) do { if (adapt) WGPP_LAUNCH(true, W); else WGPP_LAUNCH(false, W); } while (0)
    switch (tw) {
    case 8: WGPP_WIDTH(8); break;
    case 16: WGPP_WIDTH(16); break;
    case 24: WGPP_WIDTH(24); break;
    case 32: WGPP_WIDTH(32); break;
    default: return false;
    }
#undef WGPP_WIDTH
#undef WGPP_LAUNCH
    CK(cudaGetLastError());
    return true;
}

}  // namespace admmb
