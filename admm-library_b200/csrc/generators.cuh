// generators.cuh -- on-device stage-matrix generators: the step immediately BEFORE the hot path.
// A per-problem model costs 21.6 KB of H2D per problem at N = 50 (354 MB for config 4's 16,384 problems); generated from
// (e, theta0) it costs 16 B.  The kernels write straight into the batch-interleaved raw-model arrays the factor kernels
// read (rawA [36 N][ld], rawB [18 N][ld]; a shared model is [36 N], [18 N]).
//
// Oracle: oracle/gen_ocp.py spells out the same IEEE operations in the same order (library built with -fmad=false;
// sine / cosine are the polynomials below, not a libm / CUDA math-library call), so the outputs are compared bit for
// bit (tests/test_gpu_units.py::test_generator_*).  The reference has no generator (/root/reference/README.md:1-2).
#pragma once
#include "common.cuh"

namespace admmb {

constexpr int GEN_CW_IMPULSIVE = 1, GEN_CW_ZOH = 2, GEN_ELLIPTIC_ZOH = 3;

// oracle/gen_ocp.py det_sincos: Cody-Waite reduction by pi/2 in four parts, Horner polynomials on [-pi/4, pi/4]
__device__ __forceinline__ void det_sincos(double x, double &s, double &c)
{
    const double kf = rint(x * 6.36619772367581382433e-01);
    double r = x - kf * 1.57079632673412561417e+00;
    r = r - kf * 6.07710050630396597660e-11;
    r = r - kf * 2.02226624871116645580e-21;
    r = r - kf * 8.47842766036889956997e-32;
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = ps * z + -2.50507602534068634195e-08;
    ps = ps * z + 2.75573137070700676789e-06;
    ps = ps * z + -1.98412698298579493134e-04;
    ps = ps * z + 8.33333333332248946124e-03;
    ps = ps * z + -1.66666666666666324348e-01;
    const double sr = r + (r * z) * ps;
    double pc = -1.13596475577881948265e-11;
    pc = pc * z + 2.08757232129817482790e-09;
    pc = pc * z + -2.75573143513906633035e-07;
    pc = pc * z + 2.48015872894767294178e-05;
    pc = pc * z + -1.38888888888741095749e-03;
    pc = pc * z + 4.16666666666666019037e-02;
    const double cr = (1.0 - 0.5 * z) + (z * z) * pc;
    const int q = (int)((long long)kf & 3);
    s = q == 0 ? sr : q == 1 ? cr : q == 2 ? -sr : -cr;
    c = q == 0 ? cr : q == 1 ? -sr : q == 2 ? -cr : sr;
}

// Clohessy-Wiltshire closed forms (oracle/gen_ocp.py cw_stm / cw_zoh); P, G column-major 6x6 / 6x3
__device__ inline void cw_closed_form(int kind, double T, double n, double *P, double *G)
{
    for (int i = 0; i < 36; ++i) P[i] = 0.0;
    for (int i = 0; i < 18; ++i) G[i] = 0.0;
    const double nT = n * T;
    double s, c;
    det_sincos(nT, s, c);
    const double omc = 1.0 - c;
#define PE(i, j) P[(i) + 6 * (j)]
#define GE(i, j) G[(i) + 6 * (j)]
    PE(0, 0) = 4.0 - 3.0 * c;
    PE(1, 0) = 6.0 * (s - nT);
    PE(1, 1) = 1.0;
    PE(2, 2) = c;
    PE(0, 3) = s / n;
    PE(0, 4) = (2.0 * omc) / n;
    PE(1, 3) = -((2.0 * omc) / n);
    PE(1, 4) = (4.0 * s - 3.0 * nT) / n;
    PE(2, 5) = s / n;
    PE(3, 0) = (3.0 * n) * s;
    PE(4, 0) = -((6.0 * n) * omc);
    PE(5, 2) = -(n * s);
    PE(3, 3) = c;
    PE(3, 4) = 2.0 * s;
    PE(4, 3) = -(2.0 * s);
    PE(4, 4) = 4.0 * c - 3.0;
    PE(5, 5) = c;
    if (kind == GEN_CW_IMPULSIVE) {
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 6; ++i) GE(i, j) = PE(i, 3 + j);
    } else {
        const double n2 = n * n;
        GE(0, 0) = omc / n2;
        GE(0, 1) = (2.0 * (nT - s)) / n2;
        GE(1, 0) = -((2.0 * (nT - s)) / n2);
        GE(1, 1) = (4.0 * omc - 1.5 * (nT * nT)) / n2;
        GE(2, 2) = omc / n2;
        GE(3, 0) = s / n;
        GE(3, 1) = (2.0 * omc) / n;
        GE(4, 0) = -((2.0 * omc) / n);
        GE(4, 1) = (4.0 * s - 3.0 * nT) / n;
        GE(5, 2) = s / n;
    }
#undef PE
#undef GE
}

// shared CW model: thread t < 54 writes its entry for every stage
__global__ void k_gen_cw(int kind, int N, double T, double nmm, double *__restrict__ A, double *__restrict__ B)
{
    double P[36], G[18];
    cw_closed_form(kind, T, nmm, P, G);
    const int t = threadIdx.x;
    if (t < 36) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 36; ++i) v = i == t ? P[i] : v;
        for (int k = 0; k < N; ++k) A[36 * k + t] = v;
    } else if (t < 54) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 18; ++i) v = i == t - 36 ? G[i] : v;
        for (int k = 0; k < N; ++k) B[18 * k + t - 36] = v;
    }
}

struct EllCoef {
    double a30, wdot, tw, a41, k, w;
};
// oracle/gen_ocp.py _elliptic_coeffs
__device__ __forceinline__ EllCoef elliptic_coeffs(double th, double e, double p, double h)
{
    double st, ct;
    det_sincos(th, st, ct);
    const double one_ec = 1.0 + e * ct;
    const double r = p / one_ec;
    EllCoef c;
    c.w = h / (r * r);
    const double rdot = (e * st) / h;
    c.wdot = ((-2.0 * c.w) * rdot) / r;
    c.k = 1.0 / ((r * r) * r);
    const double ww = c.w * c.w;
    c.a30 = ww + 2.0 * c.k;
    c.a41 = ww - c.k;
    c.tw = 2.0 * c.w;
    return c;
}
// oracle/gen_ocp.py _f_col
__device__ __forceinline__ void f_col(const EllCoef &c, const double *y, int forced_row, double *dy)
{
    dy[0] = y[3];
    dy[1] = y[4];
    dy[2] = y[5];
    dy[3] = (c.a30 * y[0] + c.wdot * y[1]) + c.tw * y[4];
    dy[4] = ((-c.wdot) * y[0] + c.a41 * y[1]) + (-c.tw) * y[3];
    dy[5] = (-c.k) * y[2];
    if (forced_row == 3) dy[3] = dy[3] + 1.0;
    if (forced_row == 4) dy[4] = dy[4] + 1.0;
    if (forced_row == 5) dy[5] = dy[5] + 1.0;
}

// per-problem elliptic model: one thread per (problem, column j of [Phi | Gamma]); every thread integrates theta itself
// (identical bits in the nine threads of a problem).  blockIdx.y = column.
__global__ void __launch_bounds__(128) k_gen_elliptic(const double *__restrict__ ecc, const double *__restrict__ theta0,
                                                      int64_t batch, int N, double T, int substeps,
                                                      double *__restrict__ A, double *__restrict__ B, size_t ld)
{
    const int64_t p_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p_ >= batch) return;
    const int j = blockIdx.y;
    const int fr = j >= 6 ? j - 3 : -1;
    const double e = ecc[p_];
    double th = theta0[p_];
    const double p = 1.0 - e * e;
    const double h = sqrt(p);
    const double dt = T / (double)substeps;
    const double hdt = 0.5 * dt;
    const double dt6 = dt / 6.0;
    double *out = j < 6 ? A + (size_t)(6 * j) * ld + p_ : B + (size_t)(6 * (j - 6)) * ld + p_;
    const size_t stage_stride = (size_t)(j < 6 ? 36 : 18) * ld;
    for (int kst = 0; kst < N; ++kst) {
        double y[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] = (i == j) ? 1.0 : 0.0;
        for (int ss = 0; ss < substeps; ++ss) {
            const EllCoef c1 = elliptic_coeffs(th, e, p, h);
            const EllCoef c2 = elliptic_coeffs(th + hdt * c1.w, e, p, h);
            const EllCoef c3 = elliptic_coeffs(th + hdt * c2.w, e, p, h);
            const EllCoef c4 = elliptic_coeffs(th + dt * c3.w, e, p, h);
            double k1[6], k2[6], k3[6], k4[6], t[6];
            f_col(c1, y, fr, k1);
#pragma unroll
            for (int i = 0; i < 6; ++i) t[i] = y[i] + hdt * k1[i];
            f_col(c2, t, fr, k2);
#pragma unroll
            for (int i = 0; i < 6; ++i) t[i] = y[i] + hdt * k2[i];
            f_col(c3, t, fr, k3);
#pragma unroll
            for (int i = 0; i < 6; ++i) t[i] = y[i] + dt * k3[i];
            f_col(c4, t, fr, k4);
#pragma unroll
            for (int i = 0; i < 6; ++i) y[i] = y[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
            th = th + dt6 * (((c1.w + 2.0 * c2.w) + 2.0 * c3.w) + c4.w);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) out[(size_t)i * ld] = y[i];
        out += stage_stride;
    }
}

}  // namespace admmb
