#include "../admm-library_b200/csrc/kernels.cuh"
namespace admmb {
template __global__ void k_prox_cond_tf32<4>(int, const int *, const int *, int64_t, size_t, const double *, int, double, const float *, double *, double *, float *, float *, const DenseStep);
}
