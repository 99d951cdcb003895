// Micro-benchmark (developer tool): issue cost of shared-memory load patterns and shuffles, in SM cycles per instruction,
// for one warp and for four warps (one per SM sub-partition) of one CTA: is the LSU / MIO path shared by the sub-partitions?
// Patterns: 128-bit and 64-bit loads with 1 (broadcast), 2, 4, 8 and 32 distinct addresses per warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld2(unsigned a, double &x, double &y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ double ld1(unsigned a)
{
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a) : "memory");
    return r;
}
// mode 0: LDS.128, 1: LDS.64, 2: SHFL (64-bit value = 2 SHFL.32)
template <int MODE>
__global__ void k(double *out, long long *cyc, int groups, int stride_bytes, int iters)
{
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) ((double *)sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // `groups` distinct addresses per warp: lane l reads group l % groups
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)((lane % groups) * stride_bytes);
    double acc[16], v = lane;
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    double vv[4] = {v, v + 1, v + 2, v + 3};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) { double x, y; ld2(base + 16 * j, x, y); acc[j] += x; }
            else if (MODE == 1) { acc[j] += ld1(base + 8 * j); }
            else { vv[j & 3] = __shfl_sync(0xffffffffu, vv[j & 3], (lane + j) & 3, 4); }
        }
    }
    long long t1 = clock64();
    double sacc = vv[0] + vv[1] + vv[2] + vv[3];
    for (int j = 0; j < 16; ++j) sacc += acc[j];
    out[threadIdx.x] = sacc;
    if (lane == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}
template <int MODE>
void run(const char *name, int warps, int groups, int stride)
{
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    const int iters = 512;
    k<MODE><<<1, 32 * warps, 65536>>>(out, cyc, groups, stride, iters);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-8s warps %d groups %2d stride %4d B: %.2f cycles per instruction per warp, %.2f SM cycles per warp-instruction\n", name, warps,
           groups, stride, (double)mx / iters / 16, (double)mx / iters / 16 / warps);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int warps : {1, 4, 8}) {
        run<0>("LDS.128", warps, 1, 0);
        run<0>("LDS.128", warps, 2, 144);
        run<0>("LDS.128", warps, 4, 144);
        run<0>("LDS.128", warps, 8, 144);
        run<0>("LDS.128", warps, 32, 272);
        run<1>("LDS.64", warps, 1, 0);
        run<1>("LDS.64", warps, 2, 136);
        run<1>("LDS.64", warps, 4, 136);
        run<1>("LDS.64", warps, 8, 136);
        run<1>("LDS.64", warps, 32, 136);
        run<2>("SHFL.64", warps, 1, 0);
    }
    return 0;
}
