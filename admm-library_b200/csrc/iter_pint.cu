// iter_pint.cu -- launcher of the parallel-in-time kernel (iterate_pint.cuh) and of its table builder.
#include <cstdlib>
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_pint.cuh"
#include "iterate_launch_decl.cuh"

namespace admmb {

namespace {
constexpr size_t PT_SMEM_MAX = 227 * 1024;      // opt-in dynamic shared memory per CTA on sm_100

template <class K>
void pt_set_attr(K kern, size_t smem, int device)
{
    struct Done { const void *kern; int device; size_t smem; };
    static thread_local Done done[4];
    static thread_local int used = 0;
    for (int i = 0; i < used; ++i)
        if (done[i].kern == (const void *)kern && done[i].device == device && done[i].smem >= smem) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (used == 4) used = 0;
    done[used++] = Done{(const void *)kern, device, smem};
}
}  // namespace

// widest tile (problems per CTA, <= 32) whose z, u, d, a rows fit next to the factor and the tables; 0: not eligible
int iterate_pint_tile_width(const IterLaunchCtx &c)
{
    if (!c.fast_pattern || !c.decoupled || c.has_c || c.has_q || c.par_batched || c.rows_zu <= 0) return 0;
    for (int tw = 32; tw >= WG_MIN_TW; --tw)
        if (pint_layout(c.N, c.rows_zu, tw).total <= PT_SMEM_MAX) return tw;
    return 0;
}

size_t pint_table_size(int N) { return pint_table_doubles(N); }

void launch_pint_pack(cudaStream_t stream, int N, const double *fac_dec, double *table)
{
    k_pint_pack<<<1, 32, 0, stream>>>(N, fac_dec, table);
    CK(cudaGetLastError());
}

// shared factor only (the caller checks); c.pint_table = the tables of the CURRENT factor
bool launch_iterate_pint(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    const int tw = iterate_pint_tile_width(c);
    if (tw == 0 || !c.pint_table) return false;
    const size_t smem = pint_layout(c.N, c.rows_zu, tw).total;
    const int ntiles = (P.n_active + tw - 1) / tw;
    const int grid = ntiles < c.num_sms ? ntiles : c.num_sms;
#define PT_LAUNCH(A, T)                                                                          \
    do {                                                                                         \
        pt_set_attr(k_admm_iterate_pint<A, T>, smem, c.device);                                  \
        k_admm_iterate_pint<A, T><<<grid, PT_WARPS * 32, smem, c.stream>>>(P, c.pint_table, tw); \
    } while (0)
    if (adapt) { if (c.time_invariant) PT_LAUNCH(true, true); else PT_LAUNCH(true, false); }
    else { if (c.time_invariant) PT_LAUNCH(false, true); else PT_LAUNCH(false, false); }
#undef PT_LAUNCH
    CK(cudaGetLastError());
    return true;
}

}  // namespace admmb
