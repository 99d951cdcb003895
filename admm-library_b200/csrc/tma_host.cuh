// tma_host.cuh -- host side of the TMA tensor maps over the batch-interleaved working-set arrays ([rows][ld] FP64,
// boxes of [box rows] x [32 columns]).  cuTensorMapEncodeTiled is fetched through the runtime (no -lcuda).
#pragma once
#include <cuda.h>
#include "host_util.cuh"

namespace admmb {

typedef CUresult (*PFN_tmap_encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmap_encode tmap_encode_fn()
{
    static PFN_tmap_encode fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) throw CudaFail{cudaErrorNotSupported, "cuTensorMapEncodeTiled unavailable"};
        fn = (PFN_tmap_encode)p;
    }
    return fn;
}

inline CUtensorMap tmap_rows_f64(const double *base, uint64_t rows, uint64_t ld, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = tmap_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaFail{cudaErrorInvalidValue, "cuTensorMapEncodeTiled ([rows][ld] FP64 array) failed"};
    return m;
}

}  // namespace admmb
