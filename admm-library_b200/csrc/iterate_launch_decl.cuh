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
    int kernel;                   // KV_*: pin one kernel variant (tests); KV_AUTO = pick by working-set width
    int rows_zu;                  // rows of the compact z / u arrays
    int device;
    bool time_invariant;          // shared model whose A_k, B_k are the same for every stage (bitwise)
    const double *pint_table;     // parallel-in-time kernel: correction / transition tables of the current factor (or null)
    double *wgpp_blk = nullptr;   // streamed-record warp-group kernel: room for the tile-blocked copy of the stage records
                                  // (wgpp_block_doubles() doubles, owned by the shard), or null
};
void launch_iterate_smem(const IterLaunchCtx &c, const IterParams &P, bool adapt);
void launch_iterate_gshared(const IterLaunchCtx &c, const IterParams &P, bool adapt);
void launch_iterate_pp(const IterLaunchCtx &c, const IterParams &P, bool adapt);
// per-problem factors staged through shared memory by TMA; false when the configuration is not eligible
bool launch_iterate_pptma(const IterLaunchCtx &c, const IterParams &P, bool adapt);
// resident-tile kernel (iterate_res.cuh): iterates staged in shared memory for the whole launch; false when the
// configuration is not eligible (per-problem factor, generic block pattern, tile larger than shared memory)
bool launch_iterate_res(const IterLaunchCtx &c, const IterParams &P, bool adapt);
bool iterate_res_eligible(const IterLaunchCtx &c);
// warp-group kernel (iterate_wg.cuh): eight warps per resident tile; false when not eligible (needs the decoupled
// shared factor, the fast block pattern, no affine term / linear cost / per-problem parameters)
bool launch_iterate_wg(const IterLaunchCtx &c, const IterParams &P, bool adapt);
int iterate_wg_tile_width(const IterLaunchCtx &c);
// the same kernel for per-problem decoupled models: stage records streamed through a TMA ring; refactors: a rho change would
// need a new factor inside the kernel (adaptive rho with a quadratic cost), which this form does not do
bool launch_iterate_wgpp(const IterLaunchCtx &c, const IterParams &P, bool adapt);
size_t wgpp_block_doubles(const IterLaunchCtx &c, int n_active, int tw);
int iterate_wgpp_tile_width(const IterLaunchCtx &c, bool refactors);
// parallel-in-time kernel (iterate_pint.cuh): eight warps sweep eight chunks of stages at the same time;
// same eligibility as the warp-group kernel; FP64 but not the oracle's operation order, hence opt-in (KV_PINT) only
bool launch_iterate_pint(const IterLaunchCtx &c, const IterParams &P, bool adapt);
int iterate_pint_tile_width(const IterLaunchCtx &c);
size_t pint_table_size(int N);
void launch_pint_pack(cudaStream_t stream, int N, const double *fac_dec, double *table);
}  // namespace admmb
