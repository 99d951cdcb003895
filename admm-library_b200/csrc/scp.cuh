// scp.cuh -- sequential convex programming on the resident batch: the caller AFTER the hot path.
// The nonlinear relative dynamics are re-linearised about each problem's own reference trajectory on the device, the
// per-problem stage records go straight into the raw-model arrays the factor kernel reads (rawA [36 N][ld], rawB
// [18 N][ld], rawc [6 N][ld]), the batched ADMM kernels solve the convex subproblem, and a problem leaves the loop
// when its trajectory stops moving.  Nothing but 6 doubles per problem (s0) is uploaded, nothing is downloaded
// between passes.
//
// Model (ADMMB_SCP_NL_CIRCULAR): deputy about a chief on a circular orbit of radius R0, LVLH frame (x radial, y
// along-track, z cross-track), mean motion n, mu = n^2 R0^3, full two-body gravity, zero-order-hold thrust
// acceleration.  Its linearisation at r = 0 is the Clohessy-Wiltshire model.
//
// Oracle: oracle/scp_ocp.py spells out the same IEEE operations in the same order (library built with -fmad=false),
// so the stage records, every pass's ADMM solve and the final trajectories are compared bit for bit
// (tests/test_gpu_units.py::test_scp_*, tests/test_gpu_parity.py::test_scp_*).  The reference has no SCP loop
// (/root/reference/README.md:1-2).
#pragma once
#include "common.cuh"

namespace admmb {

struct ScpConst {
    double R0, twoR0, R0sq, n2, tn, dt, hdt, dt6;
    int substeps;
    int impulsive;       // the control is a velocity increment at the start of the stage, followed by a coast
    int elliptic;        // model ADMMB_SCP_NL_ELLIPTIC: R0 is the semi-major axis, time unit 1 / mean motion
};

struct ScpCoef {
    double rx, k, g, m;
};
// oracle/scp_ocp.py _coeffs: n^2 - mu / d^3 in the cancellation-free form q (3 + 3q + q^2) / (w^1.5 (1 + w^1.5))
__device__ __forceinline__ ScpCoef scp_coeffs(const double *s, const ScpConst &C)
{
    ScpCoef c;
    c.rx = C.R0 + s[0];
    const double q = (((C.twoR0 * s[0] + s[0] * s[0]) + s[1] * s[1]) + s[2] * s[2]) / C.R0sq;
    const double w = 1.0 + q;
    const double w32 = w * sqrt(w);
    c.k = C.n2 / w32;
    c.g = (C.n2 * (q * ((3.0 + 3.0 * q) + q * q))) / (w32 * (1.0 + w32));
    c.m = (3.0 * c.k) / (C.R0sq * w);
    return c;
}
// oracle/scp_ocp.py _f_state
__device__ __forceinline__ void scp_f_state(const ScpCoef &c, const double *s, const double *a, double tn, double *ds)
{
    ds[0] = s[3];
    ds[1] = s[4];
    ds[2] = s[5];
    ds[3] = (tn * s[4] + c.g * c.rx) + a[0];
    ds[4] = ((-tn) * s[3] + c.g * s[1]) + a[1];
    ds[5] = (-c.k) * s[2] + a[2];
}
// oracle/scp_ocp.py _jac + _f_col: dy = J(s) y (+ 1 on the forced row of a Gamma column)
__device__ __forceinline__ void scp_f_col(const ScpCoef &c, const double *s, const double *y, double tn, int forced_row,
                                          double *dy)
{
    const double j30 = c.g + c.m * (c.rx * c.rx), j31 = c.m * (c.rx * s[1]), j32 = c.m * (c.rx * s[2]);
    const double j41 = c.g + c.m * (s[1] * s[1]), j42 = c.m * (s[1] * s[2]), j52 = (-c.k) + c.m * (s[2] * s[2]);
    dy[0] = y[3];
    dy[1] = y[4];
    dy[2] = y[5];
    dy[3] = ((j30 * y[0] + j31 * y[1]) + j32 * y[2]) + tn * y[4];
    dy[4] = ((j31 * y[0] + j41 * y[1]) + j42 * y[2]) + (-tn) * y[3];
    dy[5] = (j32 * y[0] + j42 * y[1]) + j52 * y[2];
    if (forced_row == 3) dy[3] = dy[3] + 1.0;
    if (forced_row == 4) dy[4] = dy[4] + 1.0;
    if (forced_row == 5) dy[5] = dy[5] + 1.0;
}

// One stage about (sr, ar): F = RK4 map, A = dF/ds, B = dF/da, c = F - A sr - B ar (oracle/scp_ocp.py linearise_stage).
// The nine columns of [A | B] are integrated one after the other, each next to its own copy of the state (identical
// bits every time), so that a thread holds 12 ODE variables at a time.  Writes the stage record of problem column p.
__device__ inline void scp_linearise_stage(const ScpConst &C, const double *sr, const double *ar, double *F,
                                           double *__restrict__ Ak, double *__restrict__ Bk, double *__restrict__ ck,
                                           size_t ld)
{
    double cacc[6];
    // impulsive control (oracle/scp_ocp.py _linearise_stage_impulsive): s+ = F(s + [0; dv]) with F the coast, so A = dF/ds at the
    // post-impulse state, B = A[:, 3:6] (six columns instead of nine), c = F - A sr - B dv
    const bool imp = C.impulsive != 0;
    const double zero3[3] = {0.0, 0.0, 0.0};
    const double *af = imp ? zero3 : ar;                 // the acceleration the RK4 sees
    double s_in[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s_in[i] = (imp && i >= 3) ? sr[i] + ar[i - 3] : sr[i];
    const int ncol = imp ? 6 : 9;
    for (int j = 0; j < ncol; ++j) {
        const int fr = j >= 6 ? j - 3 : -1;
        double s[6], y[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { s[i] = s_in[i]; y[i] = (i == j) ? 1.0 : 0.0; }
        for (int ss = 0; ss < C.substeps; ++ss) {
            double k[6], l[6], ts[6], ty[6], as[6], ay[6];
            ScpCoef cf = scp_coeffs(s, C);
            scp_f_state(cf, s, af, C.tn, k);
            scp_f_col(cf, s, y, C.tn, fr, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = k[i]; ay[i] = l[i]; ts[i] = s[i] + C.hdt * k[i]; ty[i] = y[i] + C.hdt * l[i]; }
            cf = scp_coeffs(ts, C);
            scp_f_state(cf, ts, af, C.tn, k);
            scp_f_col(cf, ts, ty, C.tn, fr, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = as[i] + 2.0 * k[i]; ay[i] = ay[i] + 2.0 * l[i]; ts[i] = s[i] + C.hdt * k[i]; ty[i] = y[i] + C.hdt * l[i]; }
            cf = scp_coeffs(ts, C);
            scp_f_state(cf, ts, af, C.tn, k);
            scp_f_col(cf, ts, ty, C.tn, fr, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = as[i] + 2.0 * k[i]; ay[i] = ay[i] + 2.0 * l[i]; ts[i] = s[i] + C.dt * k[i]; ty[i] = y[i] + C.dt * l[i]; }
            cf = scp_coeffs(ts, C);
            scp_f_state(cf, ts, af, C.tn, k);
            scp_f_col(cf, ts, ty, C.tn, fr, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { s[i] = s[i] + C.dt6 * (as[i] + k[i]); y[i] = y[i] + C.dt6 * (ay[i] + l[i]); }
        }
        const double v = j < 6 ? sr[j] : ar[j - 6];
        double *out = j < 6 ? Ak + (size_t)(6 * j) * ld : Bk + (size_t)(6 * (j - 6)) * ld;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (j == 0) { cacc[i] = s[i]; F[i] = s[i]; }
            out[(size_t)i * ld] = y[i];
            cacc[i] = cacc[i] - y[i] * v;
        }
        if (imp && j >= 3) {                           // the same column is B's column j - 3
            double *outb = Bk + (size_t)(6 * (j - 3)) * ld;
            const double vb = ar[j - 3];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                outb[(size_t)i * ld] = y[i];
                cacc[i] = cacc[i] - y[i] * vb;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) ck[(size_t)i * ld] = cacc[i];
}

// ---- model ADMMB_SCP_NL_ELLIPTIC (oracle/scp_ocp.py _ell_*): the chief is on a Kepler orbit of eccentricity e; the LVLH
// frame's rotation rate w, its derivative wd, K = mu / R^3 and the chief radius R follow the true anomaly, which is
// integrated by the same RK4 (theta' = w).  Linearised at r = 0 this is config 4's model (generators.cuh k_gen_elliptic).
struct EllOrb {
    double w, wd, K, R;
};
__device__ __forceinline__ EllOrb scp_ell_coef(double th, double e, double p, double h, double a)
{
    double st, ct;
    det_sincos(th, st, ct);
    const double one_ec = 1.0 + e * ct;
    const double r = p / one_ec;
    EllOrb o;
    o.w = h / (r * r);
    const double rdot = (e * st) / h;
    o.wd = ((-2.0 * o.w) * rdot) / r;
    o.K = 1.0 / ((r * r) * r);
    o.R = a * r;
    return o;
}
struct EllGrav {
    double rx, Kg, m;
};
__device__ __forceinline__ EllGrav scp_ell_grav(const EllOrb &o, const double *s)
{
    EllGrav g;
    g.rx = o.R + s[0];
    const double q = ((((2.0 * o.R) * s[0] + s[0] * s[0]) + s[1] * s[1]) + s[2] * s[2]) / (o.R * o.R);
    const double w_ = 1.0 + q;
    const double w32 = w_ * sqrt(w_);
    const double gf = (q * ((3.0 + 3.0 * q) + q * q)) / (w32 * (1.0 + w32));
    g.Kg = o.K * gf;
    g.m = (3.0 * o.K) / (((o.R * o.R) * w_) * w32);
    return g;
}
__device__ __forceinline__ void scp_ell_f(const EllOrb &o, const EllGrav &g, const double *s, const double *a, const double *y,
                                          int forced_row, double *ds, double *dy)
{
    const double tw = 2.0 * o.w;
    const double wK = o.w * o.w - o.K;
    ds[0] = s[3];
    ds[1] = s[4];
    ds[2] = s[5];
    ds[3] = ((((tw * s[4]) + (o.wd * s[1])) + (wK * s[0])) + (g.Kg * g.rx)) + a[0];
    ds[4] = (((((-tw) * s[3]) + ((-o.wd) * s[0])) + (wK * s[1])) + (g.Kg * s[1])) + a[1];
    ds[5] = (((-o.K) * s[2]) + (g.Kg * s[2])) + a[2];
    const double j30 = (wK + g.Kg) + g.m * (g.rx * g.rx), j31 = o.wd + g.m * (g.rx * s[1]), j32 = g.m * (g.rx * s[2]);
    const double j40 = (-o.wd) + g.m * (g.rx * s[1]), j41 = (wK + g.Kg) + g.m * (s[1] * s[1]), j42 = g.m * (s[1] * s[2]);
    const double j50 = g.m * (g.rx * s[2]), j51 = g.m * (s[1] * s[2]), j52 = ((-o.K) + g.Kg) + g.m * (s[2] * s[2]);
    dy[0] = y[3];
    dy[1] = y[4];
    dy[2] = y[5];
    dy[3] = (((j30 * y[0]) + (j31 * y[1])) + (j32 * y[2])) + tw * y[4];
    dy[4] = (((j40 * y[0]) + (j41 * y[1])) + (j42 * y[2])) + (-tw) * y[3];
    dy[5] = ((j50 * y[0]) + (j51 * y[1])) + (j52 * y[2]);
    if (forced_row == 3) dy[3] = dy[3] + 1.0;
    if (forced_row == 4) dy[4] = dy[4] + 1.0;
    if (forced_row == 5) dy[5] = dy[5] + 1.0;
}

// oracle/scp_ocp.py linearise_stage_elliptic; th_k: true anomaly at the start of the stage, e: this problem's eccentricity
__device__ inline void scp_linearise_stage_ell(const ScpConst &C, double e, double th_k, const double *sr, const double *ar,
                                               double *F, double *__restrict__ Ak, double *__restrict__ Bk,
                                               double *__restrict__ ck, size_t ld)
{
    double cacc[6];
    const double p = 1.0 - e * e;
    const double h = sqrt(p);
    const bool imp = C.impulsive != 0;
    const double zero3[3] = {0.0, 0.0, 0.0};
    const double *af = imp ? zero3 : ar;
    double s_in[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s_in[i] = (imp && i >= 3) ? sr[i] + ar[i - 3] : sr[i];
    const int ncol = imp ? 6 : 9;
    for (int j = 0; j < ncol; ++j) {
        const int fr = j >= 6 ? j - 3 : -1;
        double s[6], y[6];
        double th = th_k;
#pragma unroll
        for (int i = 0; i < 6; ++i) { s[i] = s_in[i]; y[i] = (i == j) ? 1.0 : 0.0; }
        for (int ss = 0; ss < C.substeps; ++ss) {
            double k[6], l[6], ts[6], ty[6], as[6], ay[6];
            const EllOrb o1 = scp_ell_coef(th, e, p, h, C.R0);
            const EllOrb o2 = scp_ell_coef(th + C.hdt * o1.w, e, p, h, C.R0);
            const EllOrb o3 = scp_ell_coef(th + C.hdt * o2.w, e, p, h, C.R0);
            const EllOrb o4 = scp_ell_coef(th + C.dt * o3.w, e, p, h, C.R0);
            EllGrav g = scp_ell_grav(o1, s);
            scp_ell_f(o1, g, s, af, y, fr, k, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = k[i]; ay[i] = l[i]; ts[i] = s[i] + C.hdt * k[i]; ty[i] = y[i] + C.hdt * l[i]; }
            g = scp_ell_grav(o2, ts);
            scp_ell_f(o2, g, ts, af, ty, fr, k, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = as[i] + 2.0 * k[i]; ay[i] = ay[i] + 2.0 * l[i]; ts[i] = s[i] + C.hdt * k[i]; ty[i] = y[i] + C.hdt * l[i]; }
            g = scp_ell_grav(o3, ts);
            scp_ell_f(o3, g, ts, af, ty, fr, k, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { as[i] = as[i] + 2.0 * k[i]; ay[i] = ay[i] + 2.0 * l[i]; ts[i] = s[i] + C.dt * k[i]; ty[i] = y[i] + C.dt * l[i]; }
            g = scp_ell_grav(o4, ts);
            scp_ell_f(o4, g, ts, af, ty, fr, k, l);
#pragma unroll
            for (int i = 0; i < 6; ++i) { s[i] = s[i] + C.dt6 * (as[i] + k[i]); y[i] = y[i] + C.dt6 * (ay[i] + l[i]); }
            th = th + C.dt6 * (((o1.w + 2.0 * o2.w) + 2.0 * o3.w) + o4.w);
        }
        const double v = j < 6 ? sr[j] : ar[j - 6];
        double *out = j < 6 ? Ak + (size_t)(6 * j) * ld : Bk + (size_t)(6 * (j - 6)) * ld;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (j == 0) { cacc[i] = s[i]; F[i] = s[i]; }
            out[(size_t)i * ld] = y[i];
            cacc[i] = cacc[i] - y[i] * v;
        }
        if (imp && j >= 3) {
            double *outb = Bk + (size_t)(6 * (j - 3)) * ld;
            const double vb = ar[j - 3];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                outb[(size_t)i * ld] = y[i];
                cacc[i] = cacc[i] - y[i] * vb;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) ck[(size_t)i * ld] = cacc[i];
}

// true anomaly at the start of every stage, theta_tab [N][ld] (oracle/scp_ocp.py theta_table): the orbit's own clock, once per solve
__global__ void k_scp_theta(int64_t batch, int N, size_t ld, double dt, double hdt, double dt6, int substeps,
                            const double *__restrict__ ecc, const double *__restrict__ theta0, double *__restrict__ tab)
{
    const int64_t p_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p_ >= batch) return;
    const double e = ecc[p_];
    const double p = 1.0 - e * e;
    const double h = sqrt(p);
    double th = theta0[p_];
    for (int k = 0; k < N; ++k) {
        tab[(size_t)k * ld + p_] = th;
        for (int ss = 0; ss < substeps; ++ss) {
            const double w1 = scp_ell_coef(th, e, p, h, 1.0).w;
            const double w2 = scp_ell_coef(th + hdt * w1, e, p, h, 1.0).w;
            const double w3 = scp_ell_coef(th + hdt * w2, e, p, h, 1.0).w;
            const double w4 = scp_ell_coef(th + dt * w3, e, p, h, 1.0).w;
            th = th + dt6 * (((w1 + 2.0 * w2) + 2.0 * w3) + w4);
        }
    }
}

// passes >= 2: every stage of every still-moving problem about its reference trajectory xref [n][ld]; blockIdx.y = stage
__global__ void __launch_bounds__(128) k_scp_linearise(ScpConst C, int64_t batch, int N, size_t ld,
                                                       const int *__restrict__ active, const double *__restrict__ xref,
                                                       double *__restrict__ A, double *__restrict__ B, double *__restrict__ c,
                                                       const double *__restrict__ ecc, const double *__restrict__ theta_tab)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch || (active && !active[p])) return;
    const int k = blockIdx.y;
    double sr[6], ar[3], F[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sr[i] = xref[(size_t)(9 * k + i) * ld + p];
#pragma unroll
    for (int i = 0; i < 3; ++i) ar[i] = xref[(size_t)(9 * k + 6 + i) * ld + p];
    if (C.elliptic)
        scp_linearise_stage_ell(C, ecc[p], theta_tab[(size_t)k * ld + p], sr, ar, F, A + (size_t)36 * k * ld + p,
                                B + (size_t)18 * k * ld + p, c + (size_t)6 * k * ld + p, ld);
    else
        scp_linearise_stage(C, sr, ar, F, A + (size_t)36 * k * ld + p, B + (size_t)18 * k * ld + p, c + (size_t)6 * k * ld + p, ld);
}

// pass 1: the reference is the nonlinear trajectory from s0 under the controls of xref (zero: free drift), linearised
// on the way (oracle/scp_ocp.py shoot); one thread per problem, the stages follow each other
__global__ void __launch_bounds__(128) k_scp_shoot(ScpConst C, int64_t batch, int N, size_t ld, const double *__restrict__ s0,
                                                   double *__restrict__ xref, double *__restrict__ A, double *__restrict__ B,
                                                   double *__restrict__ c, const double *__restrict__ ecc,
                                                   const double *__restrict__ theta_tab)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    double sr[6], ar[3], F[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sr[i] = s0[(size_t)i * ld + p];
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = 0; i < 6; ++i) xref[(size_t)(9 * k + i) * ld + p] = sr[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) ar[i] = xref[(size_t)(9 * k + 6 + i) * ld + p];
        if (C.elliptic)
            scp_linearise_stage_ell(C, ecc[p], theta_tab[(size_t)k * ld + p], sr, ar, F, A + (size_t)36 * k * ld + p,
                                    B + (size_t)18 * k * ld + p, c + (size_t)6 * k * ld + p, ld);
        else
            scp_linearise_stage(C, sr, ar, F, A + (size_t)36 * k * ld + p, B + (size_t)18 * k * ld + p, c + (size_t)6 * k * ld + p, ld);
#pragma unroll
        for (int i = 0; i < 6; ++i) sr[i] = F[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xref[(size_t)(9 * N + i) * ld + p] = sr[i];
}

__global__ void k_scp_init(int64_t batch, size_t ld, int *active, int *passes, int *scp_status, double *step,
                           long long *iters_total)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)ld) return;
    active[p] = p < batch;
    passes[p] = 0;
    scp_status[p] = 1;
    step[p] = __longlong_as_double(0x7ff8000000000000LL);
    iters_total[p] = 0;
}

// end of a pass: how far did the trajectory move?  step = max |x - xref|, scale = max |x| (NaN if any entry is), the
// reference becomes x, and the problem leaves the loop when step <= tol_abs + tol_rel scale and its convex solve converged
// (oracle/scp_ocp.py scp_solve)
__global__ void __launch_bounds__(128) k_scp_step(int64_t batch, int n, size_t ld, int pass, double tol_abs, double tol_rel,
                                                  const double *__restrict__ x, double *__restrict__ xref,
                                                  const int *__restrict__ iters, const int *__restrict__ status, int *active,
                                                  int *passes, int *scp_status,
                                                  double *step_out, long long *iters_total, double *hist_step,
                                                  int *n_active)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch || !active[p]) return;
    double step = 0.0, scale = 0.0;
    bool nan = false;
    for (int i = 0; i < n; ++i) {
        const size_t r = (size_t)i * ld + p;
        const double xv = x[r];
        const double dv = fabs(xv - xref[r]), av = fabs(xv);
        nan = nan || dv != dv || av != av;
        step = dv > step ? dv : step;
        scale = av > scale ? av : scale;
        xref[r] = xv;
    }
    if (nan) step = scale = __longlong_as_double(0x7ff8000000000000LL);
    passes[p] = pass;
    step_out[p] = step;
    if (hist_step) hist_step[(size_t)(pass - 1) * ld + p] = step;
    iters_total[p] += (long long)iters[p];
    if (step <= tol_abs + tol_rel * scale && status[p] == ST_CONVERGED) {   // the last convex solve must itself have converged
        active[p] = 0;
        scp_status[p] = 0;
    } else {
        atomicAdd(n_active, 1);
    }
}

// warm start of the next pass: the (z, u) just computed, u with its pending scale applied (cf. k_shift_warm)
__global__ void k_scp_warm(int64_t batch, size_t ld, int rows, const double *z, const double *u, const double *usc,
                           double *z0, double *u0)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const double sc = usc[p];
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const size_t o = (size_t)r * ld + p;
        z0[o] = z[o];
        u0[o] = u[o] * sc;
    }
}

}  // namespace admmb
