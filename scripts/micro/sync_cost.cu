// Micro-benchmark (developer tool): cost in SM cycles of the intra-CTA hand-over primitives used by the warp-group kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/sync_cost scripts/micro/sync_cost.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(long long *out, int n)
{
    __shared__ __align__(16) unsigned long long bar[4];
    __shared__ uint32_t flag[4];
    __shared__ double data[32 * 8];
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(bar), fa = (uint32_t)__cvta_generic_to_shared(flag);
    const uint32_t da = (uint32_t)__cvta_generic_to_shared(data) + 8 * threadIdx.x;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba + 8));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ba + 8));   // phase 0 of bar[1] complete
        flag[0] = 1;
    }
    __syncthreads();
    long long t0, t1;
    // (a) 3 stores + release store of a flag (MEMBAR.ALL.CTA + STS)
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 256), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 512), "d"((double)i) : "memory");
        __syncwarp();
        if (threadIdx.x == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(fa + 4), "r"(i) : "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / n;
    // (a2) same with a plain (volatile) flag store
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 256), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 512), "d"((double)i) : "memory");
        __syncwarp();
        if (threadIdx.x == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(fa + 4), "r"(i) : "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / n;
    // (a3) same with an mbarrier arrive by the elected lane
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 256), "d"((double)i) : "memory");
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(da + 512), "d"((double)i) : "memory");
        __syncwarp();
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(ba) : "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / n;
    // (b) try_wait on a completed phase, dependent branch each time
    uint32_t acc = 0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ba + 8), "r"(0) : "memory");
        if (!ok) break;
        acc += ok;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / n;
    // (c) test_wait on a completed phase
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ba + 8), "r"(0) : "memory");
        if (!ok) break;
        acc += ok;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0) / n;
    // (d) acquire load poll of a flag that is already set, dependent branch each time
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        uint32_t v;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(fa) : "memory");
        if (v != 1) break;
        acc += v;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0) / n;
    // (e) dependent LDS.64 round trip (pointer-chase style)
    uint32_t a = da;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double x;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a) : "memory");
        a = da + ((uint32_t)__double2loint(x) & 0u);
    }
    t1 = clock64();
    if (threadIdx.x == 0) { out[6] = (t1 - t0) / n; out[7] = acc + a; }
}

int main()
{
    long long *d, h[8];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 32>>>(d, 2000);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[7] = {"3 STS + release flag store (MEMBAR.ALL.CTA + STS)", "3 STS + volatile flag store", "3 STS + mbarrier.arrive.release",
                            "try_wait on a completed phase + branch", "test_wait on a completed phase + branch",
                            "ld.acquire poll of a set flag + branch", "dependent LDS.64 round trip"};
    for (int i = 0; i < 7; ++i) printf("%-55s %lld cycles\n", names[i], h[i]);
    return cudaGetLastError() != cudaSuccess;
}
