// Standalone harness for the tcgen05 TF32 GEMM of csrc/dense_tf32.cuh (fast edit-compile-run loop).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include "../../admm-library_b200/csrc/dense_tf32.cuh"
using namespace admmb;
int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 456;
    const int64_t batch = argc > 2 ? atoll(argv[2]) : 256;
    const int split = argc > 3 ? atoi(argv[3]) : 1;
    const int ident = argc > 4 ? atoi(argv[4]) : 0;
    const size_t ld = round_up((size_t)batch, 32);
    std::vector<double> M((size_t)n * n), S((size_t)n * 6), mc(n), s0(6 * ld, 0.0), rt((size_t)n * ld, 0.0), ref((size_t)n * ld);
    srand(1);
    auto rnd = [] { return (rand() / (double)RAND_MAX) * 2 - 1; };
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < n; ++k) M[(size_t)i * n + k] = ident ? (i == k) : rnd() / sqrt((double)n);
        for (int j = 0; j < 6; ++j) S[(size_t)i * 6 + j] = ident ? 0 : rnd();
        mc[i] = ident ? 0 : rnd();
    }
    for (int64_t p = 0; p < batch; ++p) {
        for (int j = 0; j < 6; ++j) s0[j * ld + p] = rnd();
        for (int k = 0; k < n; ++k) rt[(size_t)k * ld + p] = rnd();
    }
    const int64_t pstep = batch > 512 ? batch / 256 : 1;      // reference on a sample of the problems
    for (int i = 0; i < n; ++i)
        for (int64_t p = 0; p < batch; p += pstep) {
            double a = mc[i];
            for (int j = 0; j < 6; ++j) a += S[(size_t)i * 6 + j] * s0[j * ld + p];
            for (int k = 0; k < n; ++k) a += M[(size_t)i * n + k] * rt[(size_t)k * ld + p];
            ref[(size_t)i * ld + p] = a;
        }
    try {
        DevBuf<double> dM, dS, dmc, ds0, drt;
        dM.alloc(M.size()); dS.alloc(S.size()); dmc.alloc(n); ds0.alloc(s0.size()); drt.alloc(rt.size());
        cudaMemcpy(dM.p, M.data(), M.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dS.p, S.data(), S.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dmc.p, mc.data(), n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(ds0.p, s0.data(), s0.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(drt.p, rt.data(), rt.size() * 8, cudaMemcpyHostToDevice);
        Tf32Plan plan;
        plan.prepare(n, batch, ld, split, dM.p, dS.p, dmc.p, ds0.p, 0);
        dim3 g((unsigned)((ld + 127) / 128), (unsigned)n);
        k_tf32_split_rows<<<g, 128>>>(n, ld, drt.p, plan.Bhi.p, split == 3 ? plan.Blo.p : nullptr);
        cudaMemset(plan.X.p, 0xff, sizeof(float) * plan.mpad * ld);   // NaN pattern: unwritten outputs show up
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        plan.gemm(0);
        cudaEventRecord(e0);
        for (int r = 0; r < 10; ++r) plan.gemm(0);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        printf("sync: %s\n", cudaGetErrorString(e));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<float> X((size_t)plan.mpad * ld);
        cudaMemcpy(X.data(), plan.X.p, X.size() * 4, cudaMemcpyDeviceToHost);
        double emax = 0, rmax = 0; long nan = 0, zero = 0;
        for (int i = 0; i < n; ++i)
            for (int64_t p = 0; p < batch; p += pstep) {
                double v = X[(size_t)i * ld + p], r = ref[(size_t)i * ld + p];
                if (v != v) { ++nan; continue; }
                if (v == 0) ++zero;
                emax = fmax(emax, fabs(v - r)); rmax = fmax(rmax, fabs(r));
            }
        printf("n %d batch %ld split %d: max err %.3e (ref max %.3e) rel %.3e, unwritten %ld, zeros %ld, %.3f ms/gemm, %.1f TFLOP/s\n", n,
               (long)batch, split, emax, rmax, emax / rmax, nan, zero, ms / 10, 2.0 * plan.mpad * plan.kpad * ld * (split == 3 ? 3 : 1) / (ms / 10 * 1e-3) / 1e12);
        for (int i = 0; i < 4; ++i) printf("  x[%d][0..3] = %g %g %g %g | ref %g %g %g %g\n", i, X[i * ld], X[i * ld + 1], X[i * ld + 2], X[i * ld + 3],
                                           ref[i * ld], ref[i * ld + 1], ref[i * ld + 2], ref[i * ld + 3]);
    } catch (const CudaFail &f) {
        printf("CUDA failure %s at %s\n", cudaGetErrorString(f.e), f.what);
        return 1;
    }
    return 0;
}
