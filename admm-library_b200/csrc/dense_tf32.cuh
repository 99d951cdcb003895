// dense_tf32.cuh -- TF32 tensor-core (tcgen05 + TMEM + TMA) version of the dense x-update.
// Placeholder wiring until the kernel lands: reports "not available" instead of silently computing
// in another precision.
#pragma once

namespace {

void dense_tf32_xupdate(Shard &)
{
    throw CudaFail{cudaErrorNotSupported, "TF32 dense x-update is not built into this library yet"};
}

int dense_tf32_unit(Shard &, int, int64_t, size_t, const double *, const double *, const double *, const double *,
                    const double *, double *)
{
    return ADMMB_E_BADARG;
}

}  // namespace
