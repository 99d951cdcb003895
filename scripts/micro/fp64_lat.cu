// Micro-benchmark (developer tool): DFMA / DADD dependent-issue latency and single-warp issue interval, in SM cycles
// (clock64 inside the kernel, one warp on one SM, so no clock-rate assumption).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps %d ILP %2d: %.2f cycles per dependent step, %.2f cycles per DFMA (one warp)\n", warps, ILP, (double)h / iters,
           (double)h / iters / ILP);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1>(1); run<2>(1); run<4>(1); run<8>(1); run<16>(1);
    run<1>(4); run<4>(4); run<8>(4); run<4>(8); run<8>(8);
    return 0;
}
