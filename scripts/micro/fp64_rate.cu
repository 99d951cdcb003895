// Microbenchmark: DFMA issue rate per SM sub-partition on the box's GPU, as a function of resident warps and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double clock_ghz)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 2048);
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<sms, warps_per_sm * 32>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)iters * ILP * warps_per_sm;                     // warp-instructions per SM
    double cyc = ms * 1e-3 * clock_ghz * 1e9;
    printf("ILP %2d warps/SM %2d: %.3f DFMA warp-instr/cycle/SM  (%.2f cycles per DFMA per SMSP-warp slot), %.1f TFLOP/s\n", ILP,
           warps_per_sm, instr / cyc, cyc / (instr / 4.0), instr * 64 * sms / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w, p.multiProcessorCount, ghz); run<4>(w, p.multiProcessorCount, ghz); run<8>(w, p.multiProcessorCount, ghz); }
    return 0;
}
