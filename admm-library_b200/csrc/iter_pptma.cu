// iter_pptma.cu -- launcher of the TMA-staged per-problem-factor kernel (iterate_pptma.cuh)
#include <cstdlib>
#include "host_util.cuh"
#define ADMMB_ITERATE_ONLY
#include "iterate_pptma.cuh"
#include "iterate_launch_decl.cuh"

namespace admmb {

namespace {
typedef CUresult (*PFN_encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encode encode_fn()
{
    static PFN_encode fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) throw CudaFail{cudaErrorNotSupported, "cuTensorMapEncodeTiled unavailable"};
        fn = (PFN_encode)p;
    }
    return fn;
}
CUtensorMap map_rows(const double *base, uint64_t rows, uint64_t ld, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaFail{cudaErrorInvalidValue, "cuTensorMapEncodeTiled (per-problem factor) failed"};
    return m;
}
}  // namespace

bool launch_iterate_pptma(const IterLaunchCtx &c, const IterParams &P, bool adapt)
{
    static const bool enabled = getenv("ADMMB_PPTMA") ? atoi(getenv("ADMMB_PPTMA")) != 0 : true;
    if (!enabled || !c.fast_pattern || !c.decoupled) return false;
    const size_t rows = (size_t)FD * c.N;
    PpTmaMaps maps;
    maps.m46 = map_rows(P.fac_dec, rows, P.ld, 46);
    maps.m10 = map_rows(P.fac_dec, rows, P.ld, 10);
    maps.m30 = map_rows(P.fac_dec, rows, P.ld, 30);
    maps.m6 = map_rows(P.fac_dec, rows, P.ld, 6);
    size_t smem = 16 + (c.par_batched ? 0 : sizeof(double) * 8 * c.nb) + sizeof(int) * ((c.nb + 3) / 4) * 4 + PPT_WARPS * 2 * 8 +
                  128 + (size_t)PPT_WARPS * 2 * ppt_slot_bytes(c.has_c);
    smem = round_up(smem, 128);
    const int T = PPT_WARPS * 32;
    const unsigned grid = (unsigned)((P.n_active + T - 1) / T);
#define PPT_LAUNCH(C, Q, A)                                                                                          \
    do {                                                                                                             \
        CK(cudaFuncSetAttribute(k_admm_iterate_pptma<C, Q, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_admm_iterate_pptma<C, Q, A><<<grid, T, smem, c.stream>>>(P, maps);                                         \
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
