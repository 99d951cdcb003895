// iter_res.cu -- launcher of the resident-tile kernel (iterate_res.cuh): tensor maps over the working-set z, u, d
// arrays, shared-memory budget, grid = one warp per 32-problem tile (persistent over tiles when there are more
// tiles than resident CTAs).
#include <cstdlib>
#include "host_util.cuh"
#include "tma_host.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_res.cuh"
#include "iterate_launch_decl.cuh"

namespace admmb {

namespace {
constexpr size_t SMEM_MAX = 227 * 1024;      // opt-in dynamic shared memory per CTA on sm_100

// rows = box * nbox with box <= 256: exact boxes, so the TMA never clips and the transaction byte count is exact
void split_rows(int rows, int &box, int &nbox)
{
    nbox = (rows + 255) / 256;
    while (rows % nbox) ++nbox;
    box = rows / nbox;
}

struct ResOcc {
    const void *kern = nullptr;
    size_t smem = 0;
    int device = -1;
    int ctas_per_sm = 0;
};

template <class K>
int res_ctas_per_sm(K kern, size_t smem, int device)
{
    static thread_local ResOcc cache[16];
    static thread_local int used = 0;
    size_t attr = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].kern == (const void *)kern && cache[i].device == device) {
            attr = attr > cache[i].smem ? attr : cache[i].smem;
            if (cache[i].smem == smem) return cache[i].ctas_per_sm;
        }
    if (smem > attr) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem));
    if (used == 16) used = 0;
    cache[used++] = ResOcc{(const void *)kern, smem, device, occ};
    return occ;
}

template <class K>
void res_launch(K kern, const IterLaunchCtx &c, const IterParams &P, const ResMaps &maps, size_t smem)
{
    const int per_sm = res_ctas_per_sm(kern, smem, c.device);
    if (per_sm < 1) throw CudaFail{cudaErrorLaunchOutOfResources, "resident-tile kernel does not fit an SM"};
    const int ntiles = (P.n_active + 31) / 32;
    const int grid = ntiles < per_sm * c.num_sms ? ntiles : per_sm * c.num_sms;
    kern<<<grid, 32, smem, c.stream>>>(P, maps);
    CK(cudaGetLastError());
}
}  // namespace

bool iterate_res_eligible(const IterLaunchCtx &c)
{
    if (!c.fast_pattern || c.rows_zu <= 0) return false;
    return res_smem_bytes(c.N, c.nb, c.rows_zu, c.decoupled, c.par_batched) <= SMEM_MAX;
}

// shared factor only (the caller checks); P.rows_zu must be set
bool launch_iterate_res(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    if (!iterate_res_eligible(c)) return false;
    ResMaps maps;
    split_rows(c.rows_zu, maps.zu_box, maps.zu_nbox);
    split_rows(3 * c.N, maps.d_box, maps.d_nbox);
    maps.z = tmap_rows_f64(P.z, (uint64_t)c.rows_zu, P.ld, (uint32_t)maps.zu_box);
    maps.u = tmap_rows_f64(P.u, (uint64_t)c.rows_zu, P.ld, (uint32_t)maps.zu_box);
    maps.d = tmap_rows_f64(P.d, (uint64_t)(3 * c.N), P.ld, (uint32_t)maps.d_box);
    const size_t smem = res_smem_bytes(c.N, c.nb, c.rows_zu, c.decoupled, c.par_batched);
#define RES_LAUNCH(C, Q, A)                                                                      \
    do {                                                                                         \
        if (c.decoupled) res_launch(k_admm_iterate_res<C, Q, A, 2>, c, P, maps, smem);           \
        else res_launch(k_admm_iterate_res<C, Q, A, 1>, c, P, maps, smem);                       \
    } while (0)
    if (c.has_c) {
        if (c.has_q) { if (adapt) RES_LAUNCH(true, true, true); else RES_LAUNCH(true, true, false); }
        else { if (adapt) RES_LAUNCH(true, false, true); else RES_LAUNCH(true, false, false); }
    } else {
        if (c.has_q) { if (adapt) RES_LAUNCH(false, true, true); else RES_LAUNCH(false, true, false); }
        else { if (adapt) RES_LAUNCH(false, false, true); else RES_LAUNCH(false, false, false); }
    }
#undef RES_LAUNCH
    return true;
}

}  // namespace admmb
