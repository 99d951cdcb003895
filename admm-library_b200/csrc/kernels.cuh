// kernels.cuh -- hand-written sm_100a kernels of the FP64 Riccati path.
//
// Data layout in HBM: batch-interleaved ("[row][ld]"): entry r of problem p lives at base[r*ld+p]
// with ld = batch rounded up to 32.  One thread owns one problem, so a warp touches 32 consecutive
// doubles (256 B) of one row per access -- fully coalesced -- and every per-problem reduction
// (the five residual norms) is thread-local.  Shared-model data (the Riccati factor, the prox
// parameter table) is staged in shared memory once per CTA and read by broadcast.
//
// Arithmetic: every operation is spelled out (mul/add/fma, fixed order) and the file is compiled
// with -fmad=false, so results are bit-identical with the canonical-order oracle used by tests/
// (SURVEY.md 7.3 H3).  Rows of SURVEY.md 8(a): a1 riccati_factor_dev / k_riccati_factor,
// a2 backward_sweep + forward part of admm_iteration / k_xupdate_riccati, a3 prox_block_dev,
// a4 process_block (relaxation, dual ascent, norms), a5 + a6 k_admm_iterate.
#pragma once
#include "common.cuh"

namespace admmb {

// ------------------------------------------------------------------------------------------------
// 3x3 Cholesky helpers (row a1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chol3_dev(const double (&H)[3][3], double (&L)[6])
{
    double l00 = sqrt(H[0][0]);
    double l10 = H[1][0] / l00;
    double l20 = H[2][0] / l00;
    double l11 = sqrt(fma(-l10, l10, H[1][1]));
    double l21 = fma(-l20, l10, H[2][1]) / l11;
    double l22 = sqrt(fma(-l21, l21, fma(-l20, l20, H[2][2])));
    L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
    return (l00 > 0.0 && l11 > 0.0 && l22 > 0.0) ? 0 : 1;
}

__device__ __forceinline__ void chol3_solve_dev(const double (&L)[6], double b0, double b1, double b2,
                                                double &x0, double &x1, double &x2)
{
    double y0 = b0 / L[0];
    double y1 = fma(-L[1], y0, b1) / L[2];
    double y2 = fma(-L[4], y1, fma(-L[3], y0, b2)) / L[5];
    x2 = y2 / L[5];
    x1 = fma(-L[4], x2, y1) / L[2];
    x0 = fma(-L[3], x2, fma(-L[1], x1, y0)) / L[0];
}

// ------------------------------------------------------------------------------------------------
// Row a1: Riccati factor of ONE model.  Raw model rows are MATLAB column-major element indices;
// element e of array X is at X[e*ldr] (pointer pre-offset by the problem index); factor entry
// (k, off) is written to fac[(k*FS+off)*ldf].
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ int riccati_factor_dev(int N, const double *A, const double *B, const double *c,
                                               const double *Q, const double *R, size_t ldr, double rho,
                                               const int *bdesc, double *fac, size_t ldf)
{
    double P[6][6], Pk[6][6], T1[6][6], BtP[3][6], G1[3][6], H[3][3], L[6];
    double Am[6][6], Bm[6][3], Km[3][6], Acl[6][6], cv[6];
    const double rinv = 1.0 / rho;
    int bad = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double w = (r == i && (bdesc[3 * N + r / 3] & 0xff) != BLK_NONE) ? 1.0 : 0.0;
            if (Q) w = fma(Q[(size_t)(36 * N + r + 6 * i) * ldr], rinv, w);
            P[r][i] = w;
        }
    for (int k = N - 1; k >= 0; --k) {
        double *f = fac + (size_t)k * FS * ldf;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) Am[i][j] = A[(size_t)(36 * k + i + 6 * j) * ldr];
#pragma unroll
            for (int j = 0; j < 3; ++j) Bm[i][j] = B[(size_t)(18 * k + i + 6 * j) * ldr];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int l = 0; l < 6; ++l) {
                double acc = Bm[0][j] * P[0][l];
#pragma unroll
                for (int i = 1; i < 6; ++i) acc = fma(Bm[i][j], P[i][l], acc);
                BtP[j][l] = acc;
            }
        const double wc = ((bdesc[3 * k + 2] & 0xff) != BLK_NONE) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int m = 0; m <= j; ++m) {
                double acc = (j == m) ? wc : 0.0;
                if (R) acc = fma(R[(size_t)(9 * k + j + 3 * m) * ldr], rinv, acc);
#pragma unroll
                for (int l = 0; l < 6; ++l) acc = fma(BtP[j][l], Bm[l][m], acc);
                H[j][m] = acc; H[m][j] = acc;
            }
        bad |= chol3_dev(H, L);
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = BtP[j][0] * Am[0][i];
#pragma unroll
                for (int l = 1; l < 6; ++l) acc = fma(BtP[j][l], Am[l][i], acc);
                G1[j][i] = acc;
            }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double x0, x1, x2;
            chol3_solve_dev(L, G1[0][i], G1[1][i], G1[2][i], x0, x1, x2);
            Km[0][i] = -x0; Km[1][i] = -x1; Km[2][i] = -x2;
            chol3_solve_dev(L, Bm[i][0], Bm[i][1], Bm[i][2], x0, x1, x2);
            f[(size_t)(F_E + 0 * 6 + i) * ldf] = x0;
            f[(size_t)(F_E + 1 * 6 + i) * ldf] = x1;
            f[(size_t)(F_E + 2 * 6 + i) * ldf] = x2;
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            double x0, x1, x2;
            chol3_solve_dev(L, m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0, x0, x1, x2);
            f[(size_t)(F_HINV + 0 * HINV_LD + m) * ldf] = x0;
            f[(size_t)(F_HINV + 1 * HINV_LD + m) * ldf] = x1;
            f[(size_t)(F_HINV + 2 * HINV_LD + m) * ldf] = x2;
        }
#pragma unroll
        for (int l = 0; l < 6; ++l)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = Am[l][i];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc = fma(Bm[l][j], Km[j][i], acc);
                Acl[l][i] = acc;
            }
#pragma unroll
        for (int r = 0; r < 6; ++r) cv[r] = c ? c[(size_t)(6 * k + r) * ldr] : 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double acc = P[r][0] * cv[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(P[r][l], cv[l], acc);
            f[(size_t)(F_CHAT + r) * ldf] = acc;
            f[(size_t)(F_C + r) * ldf] = cv[r];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) f[(size_t)(F_HINV + j * HINV_LD + 3) * ldf] = 0.0;
#pragma unroll
        for (int l = 0; l < 6; ++l) f[(size_t)(F_B + l * B_LD + 3) * ldf] = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 6; ++i) f[(size_t)(F_K + j * 6 + i) * ldf] = Km[j][i];
#pragma unroll
        for (int l = 0; l < 6; ++l) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                f[(size_t)(F_ACL + l * 6 + i) * ldf] = Acl[l][i];
                f[(size_t)(F_A + l * 6 + i) * ldf] = Am[l][i];
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) f[(size_t)(F_B + l * B_LD + j) * ldf] = Bm[l][j];
        }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = P[r][0] * Acl[0][i];
#pragma unroll
                for (int l = 1; l < 6; ++l) acc = fma(P[r][l], Acl[l][i], acc);
                T1[r][i] = acc;
            }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = (r == i && (bdesc[3 * k + r / 3] & 0xff) != BLK_NONE) ? 1.0 : 0.0;
                if (Q) acc = fma(Q[(size_t)(36 * k + r + 6 * i) * ldr], rinv, acc);
#pragma unroll
                for (int l = 0; l < 6; ++l) acc = fma(Am[l][r], T1[l][i], acc);
                Pk[r][i] = acc;
            }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int i = 0; i < 6; ++i) P[r][i] = 0.5 * (Pk[r][i] + Pk[i][r]);
    }
    return bad;
}

#ifndef ADMMB_ITERATE_ONLY
// one thread per factor.  raw_batched: model p at column p of the raw arrays (stride ld), else one
// contiguous shared model; fac_batched: factor written at column p (stride ld), else contiguous.
__global__ void k_riccati_factor(int N, int64_t batch, int raw_batched, int fac_batched, const double *A,
                                 const double *B, const double *c, const double *Q, const double *R, size_t ld,
                                 const double *rho, double rho_shared, const int *bdesc, double *fac,
                                 int *status)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const size_t ldr = raw_batched ? ld : 1, off = raw_batched ? (size_t)p : 0;
    const size_t ldf = fac_batched ? ld : 1, offf = fac_batched ? (size_t)p : 0;
    const double r = rho ? rho[p] : rho_shared;
    int bad = riccati_factor_dev(N, A + off, B + off, c ? c + off : nullptr, Q ? Q + off : nullptr,
                                 R ? R + off : nullptr, ldr, r, bdesc, fac + offf, ldf);
    if (bad && status) status[p] = ST_NAN;
}

#endif  // ADMMB_ITERATE_ONLY

// ------------------------------------------------------------------------------------------------
// Row a3: prox of one 3-block.  par(slot) reads parameter `slot` of this block.
// Every data-dependent choice is a select (setp + selp), never a branch: neighbouring problems saturate,
// shrink to zero or stay interior independently, and divergent branches here cost a latency-bound warp
// 33 % of its iteration time (measured: 1.71 ms -> 2.27 ms per 50 iterations once the lanes' active sets
// differ).  The selected values are exactly the oracle's; unselected ones (e.g. 0/0) are discarded.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sel_gt(double a, double b, double x, double y)   // (a > b) ? x : y
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %3, %4, p;\n\t}"
        : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
    return r;
}
__device__ __forceinline__ double sel_lt(double a, double b, double x, double y)   // (a < b) ? x : y
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %3, %4, p;\n\t}"
        : "=d"(r) : "d"(a), "d"(b), "d"(x), "d"(y));
    return r;
}
__device__ __forceinline__ double clip_sel(double t, double lo, double hi)          // t < lo ? lo : (t > hi ? hi : t)
{
    return sel_lt(t, lo, lo, sel_gt(t, hi, hi, t));
}

template <class ParFn>
__device__ __forceinline__ void prox_block_dev(int type, ParFn par, double rinv, const double (&v)[3],
                                               double (&z)[3])
{
    switch (type) {
    case BLK_L1:
    case BLK_L1_BOX: {
        const double kap = par(PAR_LAM) * rinv;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            // |v| - kap equals v - kap (v > kap) and -(v + kap) (v < -kap) bit for bit, so this is the
            // oracle's three-way soft threshold
            const double m = fabs(v[e]) - kap;
            double t = sel_gt(m, 0.0, copysign(m, v[e]), 0.0);
            if (type == BLK_L1_BOX) t = clip_sel(t, par(PAR_LO + e), par(PAR_HI + e));
            z[e] = t;
        }
        break;
    }
    case BLK_L2:
    case BLK_L2_BALL: {
        const double kap = par(PAR_LAM) * rinv;
        double sq = v[0] * v[0];
        sq = fma(v[1], v[1], sq);
        sq = fma(v[2], v[2], sq);
        const double nrm = sqrt(sq);
        double mag = nrm - kap;
        if (type == BLK_L2_BALL) { const double rad = par(PAR_RAD); mag = sel_gt(mag, rad, rad, mag); }
        const double sc = mag / nrm;                       // discarded when nrm <= kap (incl. 0/0)
#pragma unroll
        for (int e = 0; e < 3; ++e) z[e] = sel_gt(nrm, kap, sc * v[e], 0.0);
        break;
    }
    case BLK_BOX:
#pragma unroll
        for (int e = 0; e < 3; ++e) z[e] = clip_sel(v[e], par(PAR_LO + e), par(PAR_HI + e));
        break;
    case BLK_BALL: {
        const double c0 = par(PAR_LO), c1 = par(PAR_LO + 1), c2 = par(PAR_LO + 2), rad = par(PAR_RAD);
        const double w0 = v[0] - c0, w1 = v[1] - c1, w2 = v[2] - c2;
        double sq = w0 * w0;
        sq = fma(w1, w1, sq);
        sq = fma(w2, w2, sq);
        const double nrm = sqrt(sq);
        const double sc = rad / nrm;                       // discarded when nrm <= rad
        z[0] = sel_gt(nrm, rad, fma(sc, w0, c0), v[0]);
        z[1] = sel_gt(nrm, rad, fma(sc, w1, c1), v[1]);
        z[2] = sel_gt(nrm, rad, fma(sc, w2, c2), v[2]);
        break;
    }
    case BLK_POINT:
        z[0] = par(PAR_LO); z[1] = par(PAR_LO + 1); z[2] = par(PAR_LO + 2);
        break;
    default:
        z[0] = v[0]; z[1] = v[1]; z[2] = v[2];
        break;
    }
}

// ------------------------------------------------------------------------------------------------
// parameters of the persistent multi-iteration kernel
// ------------------------------------------------------------------------------------------------
struct IterParams {
    int N, nb, n;
    int n_active;                 // width of the working set (columns 0 .. n_active-1)
    int n_real;                   // columns n_real .. n_active-1 are whole-warp padding: copies of a running problem that
                                  // evolve identically and must not count in the statistics
    const int *orig;              // home column of each working-set column (nullptr: identity)
    size_t hist_ld;               // leading dimension of the history arrays (home layout)
    size_t ld;
    const double *fac;            // shared: [FS*N]; per problem: [FS*N][ld]
    double *fac_rw;               // per-problem factor, writable (refactor)
    const double *fac_dec;        // packed decoupled records, shared [FD*N] or per problem [FD*N][ld]
    double *fac_dec_rw;
    const double *rawA, *rawB, *rawc, *rawQ, *rawR;   // raw model for the in-kernel refactor
    int raw_batched;
    const double *s0;             // [6][ld]
    double *z, *u;                // [3*nsplitblk][ld] compact rows of the split blocks
    double *d;                    // [3N][ld]: d_k of the last backward sweep (k_output rebuilds a_k, x from it)
    const double *q;              // [n] or [n][ld]
    int q_batched;
    const int *bdesc;             // [nb]: type | slot << 8 (slot = compact index of a split block)
    const double *par;            // [8*nb] or [8*nb][ld]
    int par_batched;
    double *rho, *usc;            // [ld] per-problem rho and pending dual scale
    int *iters, *status;          // [ld]
    double *fin;                  // [4][ld] r_norm, s_norm, eps_pri, eps_dual of the last iteration
    double *hist;                 // 5 x [max_iter][ld] or nullptr
    size_t hist_stride;
    unsigned long long *refac_count;
    // finished lanes keep iterating until the launch ends (see k_admm_iterate): at the moment a problem finishes its
    // z, u, d columns are copied to their home columns and snap[p] is set so that the repack does not retire them again
    double *z_home, *u_home, *d_home;   // nullptr: working set == home arrays (finished lanes leave the loop instead)
    size_t home_ld;
    int rows_zu;
    int *snap;
    double alpha, oma, reltol, sqrtn_abs, mu, tau, inv_tau;
    int adapt, every, until, max_iter, chunk, has_P;
};

// factor accessor: shared (contiguous, smem or global) or per problem (interleaved, pre-offset by p)
template <bool SHARED>
struct FacRef {
    const double *base;
    size_t ld;
    uint32_t sbase;               // shared-window address of the staged factor (0 when not in smem); for per-problem
                                  // factors staged by TMA: address of the current ring slot + 8 * lane
    __device__ __forceinline__ double operator()(int k, int off) const
    {
        if (SHARED) return base[k * FS + off];
        return base[((size_t)k * FS + off) * ld];
    }
};

// Where an iteration finds its iterates.  GlobalIO: the batch-interleaved working set in global memory, one column
// per problem (pitch P.ld, L2-cached .cg loads / stores).  SmemIO (iterate_res.cuh): the 32-problem tile of the warp,
// staged in shared memory for the whole launch ([row][32 lanes]).
struct GlobalIO {
    __device__ __forceinline__ double *z(const IterParams &P, size_t p) const { return P.z + p; }
    __device__ __forceinline__ double *u(const IterParams &P, size_t p) const { return P.u + p; }
    __device__ __forceinline__ double *d(const IterParams &P, size_t p) const { return P.d + p; }
    __device__ __forceinline__ size_t pitch(const IterParams &P) const { return P.ld; }
    // initial state s0[i] of problem p
    __device__ __forceinline__ double s0(const IterParams &P, size_t p, int i) const { return P.s0[p + (size_t)i * P.ld]; }
    static __device__ __forceinline__ double ld(const double *a) { return ADMMB_LD(a); }
    static __device__ __forceinline__ void st(double *a, double v) { ADMMB_ST(a, v); }
};

__device__ __forceinline__ double2 lds128(uint32_t a)
{
    double2 r;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ double lds64(uint32_t a)
{
    double r;
    asm("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
}
// A read of a ring slot that TMA refills (iterate_pptma.cuh): `volatile` + the memory clobber keep it between the mbarrier
// wait that publishes the slot and the __syncwarp that releases it.  A plain asm load has no side effects and depends
// only on its address, which is known before the wait -- the compiler may hoist it above the wait and read the slot's
// previous contents (seen with the two-slot ring of the generic-record kernel: NaN at > 4,736 problems).
__device__ __forceinline__ double lds64_ordered(uint32_t a)
{
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a) : "memory");
    return r;
}

// ------------------------------------------------------------------------------------------------
// One ADMM iteration of problem p (rows a2 + a3 + a4 fused): backward sweep, then a forward sweep
// that turns each stage's x into z, u and the norm accumulators without x ever leaving registers.
// ------------------------------------------------------------------------------------------------
template <bool FSH, bool HAS_C, bool HAS_Q, bool ADAPT>
__device__ __forceinline__ void admm_iteration(const IterParams &P, const size_t p, const FacRef<FSH> F,
                                               const int *bdesc, const double *parS, const double rho,
                                               const double sigma, double (&nr)[5])
{
    const int N = P.N;
    const size_t ld = P.ld;
    const double rinv = 1.0 / rho;
    const double *zp = P.z + p, *up = P.u + p;
    const double *qp = HAS_Q ? (P.q_batched ? P.q + p : P.q) : nullptr;
    const size_t qld = P.q_batched ? ld : 1;

    // right-hand side of block b: w*(z - u) - q/rho
    auto load_rt = [&](int b, double (&t)[3]) {
        const int de = bdesc[b];
        if ((de & 0xff) != BLK_NONE) {
            const size_t r0 = (size_t)(de >> 8) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double zz = ld_stream(zp + (r0 + e) * ld);
                double uu = ld_stream(up + (r0 + e) * ld);
                if (ADAPT) uu = uu * sigma;
                double v = zz - uu;
                if (HAS_Q) v = fma(-qp[(size_t)(3 * b + e) * qld], rinv, v);
                t[e] = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) t[e] = HAS_Q ? -(qp[(size_t)(3 * b + e) * qld] * rinv) : 0.0;
        }
    };

    // ---------------- backward sweep (row a2, first half)
    double g[6];
    {
        double t0[3], t1[3];
        load_rt(3 * N, t0);
        load_rt(3 * N + 1, t1);
        g[0] = t0[0]; g[1] = t0[1]; g[2] = t0[2]; g[3] = t1[0]; g[4] = t1[1]; g[5] = t1[2];
    }
    for (int k = N - 1; k >= 0; --k) {
        double rs[6], ra[3], pn[6];
        {
            double t0[3], t1[3];
            load_rt(3 * k, t0);
            load_rt(3 * k + 1, t1);
            load_rt(3 * k + 2, ra);
            rs[0] = t0[0]; rs[1] = t0[1]; rs[2] = t0[2]; rs[3] = t1[0]; rs[4] = t1[1]; rs[5] = t1[2];
        }
        if (HAS_C) {
#pragma unroll
            for (int i = 0; i < 6; ++i) g[i] = g[i] - F(k, F_CHAT + i);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = F(k, F_HINV + HINV_LD * j + 0) * ra[0];
            acc = fma(F(k, F_HINV + HINV_LD * j + 1), ra[1], acc);
            acc = fma(F(k, F_HINV + HINV_LD * j + 2), ra[2], acc);
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_E + 6 * j + i), g[i], acc);
            st_stream(P.d + p + (size_t)(3 * k + j) * ld, acc);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = rs[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_K + 6 * j + i), ra[j], acc);
#pragma unroll
            for (int l = 0; l < 6; ++l) acc = fma(F(k, F_ACL + 6 * l + i), g[l], acc);
            pn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = pn[i];
    }

    // ---------------- forward sweep fused with prox / dual ascent / norms (rows a2, a3, a4)
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    double *zw = P.z + p, *uw = P.u + p;
    auto process_block = [&](int b, const double x0, const double x1, const double x2) {
        const int de = bdesc[b];
        const int type = de & 0xff;
        if (type == BLK_NONE) return;
        const size_t r0 = (size_t)(de >> 8) * 3;
        const double xb[3] = {x0, x1, x2};
        double zo[3], v[3], zn[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            zo[e] = ld_stream(zw + (r0 + e) * ld);
            double uo = ld_stream(uw + (r0 + e) * ld);
            if (ADAPT) uo = uo * sigma;
            double xh = fma(P.alpha, xb[e], P.oma * zo[e]);
            v[e] = xh + uo;
        }
        if (P.par_batched) {
            const double *pp = P.par + p + (size_t)(8 * b) * ld;
            prox_block_dev(type, [&](int s) { return pp[(size_t)s * ld]; }, rinv, v, zn);
        } else {
            const double *pp = parS + 8 * b;
            prox_block_dev(type, [&](int s) { return pp[s]; }, rinv, v, zn);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            double un = v[e] - zn[e];
            double dr = xb[e] - zn[e];
            double ds = zn[e] - zo[e];
            rr = fma(dr, dr, rr);
            ss = fma(ds, ds, ss);
            xx = fma(xb[e], xb[e], xx);
            zz = fma(zn[e], zn[e], zz);
            uu = fma(un, un, uu);
            st_stream(zw + (r0 + e) * ld, zn[e]);
            st_stream(uw + (r0 + e) * ld, un);
        }
    };

    double s[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = P.s0[p + (size_t)i * ld];
    for (int k = 0; k < N; ++k) {
        double a[3], sn[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = ld_stream(P.d + p + (size_t)(3 * k + j) * ld);
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_K + 6 * j + i), s[i], acc);
            a[j] = acc;
        }
        process_block(3 * k, s[0], s[1], s[2]);
        process_block(3 * k + 1, s[3], s[4], s[5]);
        process_block(3 * k + 2, a[0], a[1], a[2]);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = F(k, F_A + 6 * i + 0) * s[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(F(k, F_A + 6 * i + l), s[l], acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_B + B_LD * i + j), a[j], acc);
            if (HAS_C) acc = acc + F(k, F_C + i);
            sn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = sn[i];
    }
    process_block(3 * N, s[0], s[1], s[2]);
    process_block(3 * N + 1, s[3], s[4], s[5]);
    nr[0] = rr; nr[1] = ss; nr[2] = xx; nr[3] = zz; nr[4] = uu;
}

// ------------------------------------------------------------------------------------------------
// Fast path of the same iteration for the pattern every benchmark configuration has: state blocks of
// stages 0..N-1 unsplit (BLK_NONE), every control block split, terminal blocks arbitrary.
//   * the next stage's z, u (and d) are prefetched into registers while the current stage computes,
//     so global-memory latency is off the sequential stage chain;
//   * factor rows are read with 128-bit loads (LDS.128 when the factor is staged in shared memory).
// Operation order per accumulator is exactly that of admm_iteration / the oracle.
// ------------------------------------------------------------------------------------------------
// hooks of the per-stage factor staging; the default does nothing (factor already addressable)
struct NoStaging {
    template <class FR> __device__ __forceinline__ void iter_begin(FR &) {}
    template <class FR> __device__ __forceinline__ void bwd_begin(int, FR &) {}
    template <class FR> __device__ __forceinline__ void fwd_begin(int, FR &) {}
    __device__ __forceinline__ void stage_end() {}
};

// Per-problem factors (FSH = false) with FSMEM = true: the current stage's record rows of the warp's 32 problems have
// been staged in shared memory by TMA (iterate_pptma.cuh, generic records): [row][32 lanes] doubles, F.sbase = slot + 8 *
// lane.  A backward slot holds record rows 0..83 (K, Acl, Hinv, E) followed by chat, a forward slot K followed by rows
// 84..149 (A, B, c): in both, record offset `off` sits in slot row off (off < 84) or off - 66.
__device__ __forceinline__ uint32_t staged_row(int off) { return (uint32_t)(off < F_A ? off : off - 66) * 256u; }

template <bool FSH, bool FSMEM>
__device__ __forceinline__ void fac_row6(const FacRef<FSH> &F, int k, int off, double (&r)[6])
{
    if (FSH && FSMEM) {
        const uint32_t a = F.sbase + (uint32_t)(k * FS + off) * 8u;
        const double2 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32);
        r[0] = x.x; r[1] = x.y; r[2] = y.x; r[3] = y.y; r[4] = z.x; r[5] = z.y;
    } else if (FSH) {
        const double2 *q = reinterpret_cast<const double2 *>(F.base + k * FS + off);
        const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = c.x; r[5] = c.y;
    } else if (FSMEM) {
#pragma unroll
        for (int i = 0; i < 6; ++i) r[i] = lds64_ordered(F.sbase + staged_row(off + i));
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) r[i] = F.base[((size_t)k * FS + off + i) * F.ld];
    }
}
template <bool FSH, bool FSMEM>
__device__ __forceinline__ void fac_row3(const FacRef<FSH> &F, int k, int off, double (&r)[3])
{
    if (FSH && FSMEM) {   // rows of Hinv and B are padded to 4 doubles
        const uint32_t a = F.sbase + (uint32_t)(k * FS + off) * 8u;
        const double2 x = lds128(a), y = lds128(a + 16);
        r[0] = x.x; r[1] = x.y; r[2] = y.x;
    } else if (FSH) {
        const double2 *q = reinterpret_cast<const double2 *>(F.base + k * FS + off);
        const double2 a = __ldg(q), b = __ldg(q + 1);
        r[0] = a.x; r[1] = a.y; r[2] = b.x;
    } else if (FSMEM) {
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = lds64_ordered(F.sbase + staged_row(off + i));
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = F.base[((size_t)k * FS + off + i) * F.ld];
    }
}

// relaxation + prox + dual ascent + norm accumulation of one split block whose old z, u (u already
// scaled by the pending sigma) are in registers; stores the new z, u.
template <class IO = GlobalIO, class ParFn>
__device__ __forceinline__ void block_update(int type, ParFn par, double rinv, double alpha, double oma,
                                             const double (&xb)[3], const double (&zo)[3], const double (&uo)[3],
                                             double *zrow, double *urow, size_t ld, double &rr, double &ss,
                                             double &xx, double &zz, double &uu, const bool store = true)
{
    double v[3], zn[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        double xh = fma(alpha, xb[e], oma * zo[e]);
        v[e] = xh + uo[e];
    }
    prox_block_dev(type, par, rinv, v, zn);
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        double un = v[e] - zn[e];
        double dr = xb[e] - zn[e];
        double ds = zn[e] - zo[e];
        rr = fma(dr, dr, rr);
        ss = fma(ds, ds, ss);
        xx = fma(xb[e], xb[e], xx);
        zz = fma(zn[e], zn[e], zz);
        uu = fma(un, un, uu);
        if (store) {
            IO::st(zrow + (size_t)e * ld, zn[e]);
            IO::st(urow + (size_t)e * ld, un);
        }
    }
}

template <bool FSH, bool FSMEM, bool HAS_C, bool HAS_Q, bool ADAPT, class IO = GlobalIO, class Staging = NoStaging>
__device__ __forceinline__ void admm_iteration_fast(const IterParams &P, const size_t p, FacRef<FSH> F,
                                                    const int *bdesc, const uint32_t par_sbase, const double rho,
                                                    const double sigma, double (&nr)[5], const IO io = IO(),
                                                    const bool act = true, Staging stg = Staging())
{
    const int N = P.N;
    stg.iter_begin(F);
    const size_t ld = P.ld;
    const double rinv = 1.0 / rho;
    // in this pattern the split blocks are ctrl_0 .. ctrl_{N-1} followed by the terminal blocks, so the
    // compact z/u rows of ctrl_k are 3k..3k+2: plain pointer walks, no index arithmetic in the loops
    const size_t ldi = io.pitch(P);                   // pitch of the iterate rows (z, u, d)
    const ptrdiff_t ld1 = (ptrdiff_t)ldi, ld2 = 2 * (ptrdiff_t)ldi, ld3 = 3 * (ptrdiff_t)ldi;
    double *const zp = io.z(P, p), *const up = io.u(P, p), *const dp = io.d(P, p);
    const double *qp = HAS_Q ? (P.q_batched ? P.q + p : P.q) : nullptr;
    const size_t qld = P.q_batched ? ld : 1;

    // parameters of block b into registers (shared table: four 128-bit shared loads)
    auto load_par = [&](int b, double (&pr)[8]) {
        if (P.par_batched) {
            const double *pp = P.par + p + (size_t)(8 * b) * ld;
#pragma unroll
            for (int q = 0; q < 8; ++q) pr[q] = pp[(size_t)q * ld];
        } else {
            const uint32_t a = par_sbase + (uint32_t)b * 64u;
            const double2 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32), w = lds128(a + 48);
            pr[0] = x.x; pr[1] = x.y; pr[2] = y.x; pr[3] = y.y; pr[4] = z.x; pr[5] = z.y; pr[6] = w.x; pr[7] = w.y;
        }
    };
    auto rt_terminal = [&](int b, double (&t)[3]) {
        const int de = bdesc[b];
        if ((de & 0xff) != BLK_NONE) {
            const size_t r0 = (size_t)(de >> 8) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double uu = IO::ld(up + (r0 + e) * ldi);
                if (ADAPT) uu = uu * sigma;
                double v = IO::ld(zp + (r0 + e) * ldi) - uu;
                if (HAS_Q) v = fma(-qp[(size_t)(3 * b + e) * qld], rinv, v);
                t[e] = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) t[e] = HAS_Q ? -(qp[(size_t)(3 * b + e) * qld] * rinv) : 0.0;
        }
    };

    // ---------------- backward sweep: two register buffers, prefetch distance two stages
    double gA[6], gB[6];
    {
        double t0[3], t1[3];
        rt_terminal(3 * N, t0);
        rt_terminal(3 * N + 1, t1);
        gA[0] = t0[0]; gA[1] = t0[1]; gA[2] = t0[2]; gA[3] = t1[0]; gA[4] = t1[1]; gA[5] = t1[2];
    }
    double zA[3], uA[3], zB[3], uB[3];
    {
        const double *z0 = zp + (ptrdiff_t)(N - 1) * ld3, *u0 = up + (ptrdiff_t)(N - 1) * ld3;
        zA[0] = IO::ld(z0); zA[1] = IO::ld(z0 + ld1); zA[2] = IO::ld(z0 + ld2);
        uA[0] = IO::ld(u0); uA[1] = IO::ld(u0 + ld1); uA[2] = IO::ld(u0 + ld2);
        if (N > 1) {
            z0 -= ld3; u0 -= ld3;
            zB[0] = IO::ld(z0); zB[1] = IO::ld(z0 + ld1); zB[2] = IO::ld(z0 + ld2);
            uB[0] = IO::ld(u0); uB[1] = IO::ld(u0 + ld1); uB[2] = IO::ld(u0 + ld2);
        }
    }
    const double *zl = zp + (ptrdiff_t)(N - 3) * ld3, *ul = up + (ptrdiff_t)(N - 3) * ld3;   // rows of stage k-2
    double *ds = dp + (ptrdiff_t)(N - 1) * ld3;                                               // rows of stage k
    auto bwd_stage = [&](const int k, double (&zc)[3], double (&uc)[3], const double (&g)[6], double (&pn)[6]) {
        stg.bwd_begin(k, F);
        double ra[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            double uu = ADAPT ? uc[e] * sigma : uc[e];
            double v = zc[e] - uu;
            if (HAS_Q) v = fma(-qp[(size_t)(9 * k + 6 + e) * qld], rinv, v);
            ra[e] = v;
        }
        if (k >= 2) {   // refill this buffer with the stage two steps ahead
            zc[0] = IO::ld(zl); zc[1] = IO::ld(zl + ld1); zc[2] = IO::ld(zl + ld2);
            uc[0] = IO::ld(ul); uc[1] = IO::ld(ul + ld1); uc[2] = IO::ld(ul + ld2);
        }
        zl -= ld3; ul -= ld3;
        double gg[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) gg[i] = g[i];
        if (HAS_C) {
            double ch[6];
            fac_row6<FSH, FSMEM>(F, k, F_CHAT, ch);
#pragma unroll
            for (int i = 0; i < 6; ++i) gg[i] = gg[i] - ch[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) pn[i] = HAS_Q ? -(qp[(size_t)(9 * k + i) * qld] * rinv) : 0.0;
        double dj[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double h[3], er[6];
            fac_row3<FSH, FSMEM>(F, k, F_HINV + HINV_LD * j, h);
            fac_row6<FSH, FSMEM>(F, k, F_E + 6 * j, er);
            double acc = h[0] * ra[0];
            acc = fma(h[1], ra[1], acc);
            acc = fma(h[2], ra[2], acc);
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(er[i], gg[i], acc);
            dj[j] = acc;
        }
        if (act) { IO::st(ds, dj[0]); IO::st(ds + ld1, dj[1]); IO::st(ds + ld2, dj[2]); }
        ds -= ld3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double kr[6];
            fac_row6<FSH, FSMEM>(F, k, F_K + 6 * j, kr);
#pragma unroll
            for (int i = 0; i < 6; ++i) pn[i] = fma(kr[i], ra[j], pn[i]);
        }
#pragma unroll
        for (int l = 0; l < 6; ++l) {
            double ar[6];
            fac_row6<FSH, FSMEM>(F, k, F_ACL + 6 * l, ar);
#pragma unroll
            for (int i = 0; i < 6; ++i) pn[i] = fma(ar[i], gg[l], pn[i]);
        }
        stg.stage_end();
    };
    {
        int k = N - 1;
        for (; k >= 1; k -= 2) {
            bwd_stage(k, zA, uA, gA, gB);
            bwd_stage(k - 1, zB, uB, gB, gA);
        }
        if (k == 0) bwd_stage(0, zA, uA, gA, gB);
    }

    // ---------------- forward sweep fused with prox / dual ascent / norms
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    double sA[6], sB[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sA[i] = io.s0(P, p, i);
    double dA[3], dB[3];
    {
        zA[0] = IO::ld(zp); zA[1] = IO::ld(zp + ld1); zA[2] = IO::ld(zp + ld2);
        uA[0] = IO::ld(up); uA[1] = IO::ld(up + ld1); uA[2] = IO::ld(up + ld2);
        dA[0] = IO::ld(dp); dA[1] = IO::ld(dp + ld1); dA[2] = IO::ld(dp + ld2);
        if (N > 1) {
            zB[0] = IO::ld(zp + ld3); zB[1] = IO::ld(zp + ld3 + ld1); zB[2] = IO::ld(zp + ld3 + ld2);
            uB[0] = IO::ld(up + ld3); uB[1] = IO::ld(up + ld3 + ld1); uB[2] = IO::ld(up + ld3 + ld2);
            dB[0] = IO::ld(dp + ld3); dB[1] = IO::ld(dp + ld3 + ld1); dB[2] = IO::ld(dp + ld3 + ld2);
        }
    }
    const double *zf = zp + 2 * ld3, *uf = up + 2 * ld3, *df = dp + 2 * ld3;   // rows of stage k+2 (loads)
    double *zw = zp, *uw = up;                                                  // rows of stage k (stores)
    auto fwd_stage = [&](const int k, double (&zc)[3], double (&uc)[3], double (&dc)[3], const double (&s)[6],
                         double (&sn)[6]) {
        stg.fwd_begin(k, F);
        double a[3], zo[3], uo[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double kr[6];
            fac_row6<FSH, FSMEM>(F, k, F_K + 6 * j, kr);
            double acc = dc[j];
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(kr[i], s[i], acc);
            a[j] = acc;
            zo[j] = zc[j];
            uo[j] = ADAPT ? uc[j] * sigma : uc[j];
        }
        if (k + 2 < N) {   // refill this buffer with the stage two steps ahead
            zc[0] = IO::ld(zf); zc[1] = IO::ld(zf + ld1); zc[2] = IO::ld(zf + ld2);
            uc[0] = IO::ld(uf); uc[1] = IO::ld(uf + ld1); uc[2] = IO::ld(uf + ld2);
            dc[0] = IO::ld(df); dc[1] = IO::ld(df + ld1); dc[2] = IO::ld(df + ld2);
        }
        zf += ld3; uf += ld3; df += ld3;
        {
            const int b = 3 * k + 2;
            double pr[8];
            load_par(b, pr);
            block_update<IO>(bdesc[b] & 0xff, [&](int q) { return pr[q]; }, rinv, P.alpha, P.oma, a, zo, uo, zw, uw, ldi, rr,
                         ss, xx, zz, uu, act);
            zw += ld3; uw += ld3;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double ar[6], br[3];
            fac_row6<FSH, FSMEM>(F, k, F_A + 6 * i, ar);
            fac_row3<FSH, FSMEM>(F, k, F_B + B_LD * i, br);
            double acc = ar[0] * s[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(ar[l], s[l], acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(br[j], a[j], acc);
            if (HAS_C) acc = acc + ((!FSH && FSMEM) ? lds64_ordered(F.sbase + staged_row(F_C + i)) : F(k, F_C + i));
            sn[i] = acc;
        }
        stg.stage_end();
    };
    bool s_in_A = true;
    {
        int k = 0;
        for (; k + 1 < N; k += 2) {
            fwd_stage(k, zA, uA, dA, sA, sB);
            fwd_stage(k + 1, zB, uB, dB, sB, sA);
        }
        if (k < N) { fwd_stage(k, zA, uA, dA, sA, sB); s_in_A = false; }
    }
    // terminal blocks (arbitrary types)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int b = 3 * N + t;
        const int de = bdesc[b];
        if ((de & 0xff) == BLK_NONE) continue;
        const size_t r0 = (size_t)(de >> 8) * 3;
        const double xb[3] = {s_in_A ? sA[3 * t] : sB[3 * t], s_in_A ? sA[3 * t + 1] : sB[3 * t + 1],
                              s_in_A ? sA[3 * t + 2] : sB[3 * t + 2]};
        double zo[3], uo[3], pr[8];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            zo[e] = IO::ld(zp + (r0 + e) * ldi);
            double uv = IO::ld(up + (r0 + e) * ldi);
            uo[e] = ADAPT ? uv * sigma : uv;
        }
        load_par(b, pr);
        block_update<IO>(de & 0xff, [&](int q) { return pr[q]; }, rinv, P.alpha, P.oma, xb, zo, uo, zp + r0 * ldi,
                     up + r0 * ldi, ldi, rr, ss, xx, zz, uu, act);
    }
    nr[0] = rr; nr[1] = ss; nr[2] = xx; nr[3] = zz; nr[4] = uu;
}

// ------------------------------------------------------------------------------------------------
// Decoupled models (in-plane / cross-track).  k_check_decoupled proves the structure on the computed
// factor (every off-pattern entry of K, Acl, Hinv, E, A, B is exactly zero), pack_decoupled_dev
// gathers the 88 structurally non-zero entries per stage, admm_iteration_dec is the fast iteration
// with the zero terms dropped: fma(0, x, acc) == acc, so every accumulator sees the same non-zero
// terms in the same order as in the oracle and the results stay bit-identical (signed zeros aside).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_plane_state(int i) { return i != 2 && i != 5; }

#ifndef ADMMB_ITERATE_ONLY
__global__ void k_check_decoupled(int N, int64_t batch, int fac_batched, const double *fac, size_t ld, int *flag)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const size_t ldf = fac_batched ? ld : 1;
    const double *f = fac + (fac_batched ? (size_t)p : 0);
    bool bad = false;
    auto at = [&](int k, int off) { return f[((size_t)k * FS + off) * ldf]; };
    for (int k = 0; k < N && !bad; ++k) {
        for (int j = 0; j < 3; ++j) {
            for (int i = 0; i < 6; ++i) {
                const bool on = (j < 2) == in_plane_state(i);
                if (!on) bad |= (at(k, F_K + 6 * j + i) != 0.0) || (at(k, F_E + 6 * j + i) != 0.0) ||
                                (at(k, F_B + B_LD * i + j) != 0.0);
            }
            for (int m = 0; m < 3; ++m)
                if ((j < 2) != (m < 2)) bad |= at(k, F_HINV + HINV_LD * j + m) != 0.0;
        }
        for (int l = 0; l < 6; ++l)
            for (int i = 0; i < 6; ++i)
                if (in_plane_state(l) != in_plane_state(i))
                    bad |= (at(k, F_ACL + 6 * l + i) != 0.0) || (at(k, F_A + 6 * l + i) != 0.0);
    }
    if (bad) atomicOr(flag, 1);
}

#endif  // ADMMB_ITERATE_ONLY

__device__ __forceinline__ void pack_decoupled_dev(int N, const double *src, size_t lds, double *dst, size_t ldd)
{
    const int IP[4] = {0, 1, 3, 4}, CR[2] = {2, 5};
    for (int k = 0; k < N; ++k) {
        const double *f = src + (size_t)k * FS * lds;
        double *o = dst + (size_t)k * FD * ldd;
        auto S = [&](int off) { return f[(size_t)off * lds]; };
        auto D = [&](int off, double v) { o[(size_t)off * ldd] = v; };
        for (int j = 0; j < 2; ++j)
            for (int c = 0; c < 4; ++c) {
                D(D_KIN + 4 * j + c, S(F_K + 6 * j + IP[c]));
                D(D_EIN + 4 * j + c, S(F_E + 6 * j + IP[c]));
            }
        for (int c = 0; c < 2; ++c) {
            D(D_KC + c, S(F_K + 12 + CR[c]));
            D(D_EC + c, S(F_E + 12 + CR[c]));
            D(D_BC + c, S(F_B + B_LD * CR[c] + 2));
        }
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 4; ++c) {
                D(D_ACLIN + 4 * r + c, S(F_ACL + 6 * IP[r] + IP[c]));
                D(D_AIN + 4 * r + c, S(F_A + 6 * IP[r] + IP[c]));
            }
            for (int j = 0; j < 2; ++j) D(D_BIN + 2 * r + j, S(F_B + B_LD * IP[r] + j));
        }
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 2; ++c) {
                D(D_ACLC + 2 * r + c, S(F_ACL + 6 * CR[r] + CR[c]));
                D(D_AC + 2 * r + c, S(F_A + 6 * CR[r] + CR[c]));
                D(D_HIN + 2 * r + c, S(F_HINV + HINV_LD * r + c));
            }
        D(D_HC, S(F_HINV + HINV_LD * 2 + 2));
        D(D_HC + 1, 0.0);
        for (int i = 0; i < 6; ++i) {
            D(D_C + i, S(F_C + i));
            D(D_CHAT + i, S(F_CHAT + i));
        }
    }
}

#ifndef ADMMB_ITERATE_ONLY
__global__ void k_pack_decoupled(int N, int64_t batch, int fac_batched, const double *fac, size_t ld, double *out)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const size_t l = fac_batched ? ld : 1, off = fac_batched ? (size_t)p : 0;
    pack_decoupled_dev(N, fac + off, l, out + off, l);
}

#endif  // ADMMB_ITERATE_ONLY

// W consecutive doubles of the packed record (W = 2 or 4; offsets are even => 16-byte aligned)
template <bool FSH, bool FSMEM, int W>
__device__ __forceinline__ void dec_ld(const FacRef<FSH> &F, int k, int off, double (&r)[W])
{
    if (FSH && FSMEM) {
        const uint32_t a = F.sbase + (uint32_t)(k * FD + off) * 8u;
        const double2 x = lds128(a);
        r[0] = x.x; r[1] = x.y;
        if (W == 4) { const double2 y = lds128(a + 16); r[2] = y.x; r[3] = y.y; }
    } else if (FSH) {
        const double2 *q = reinterpret_cast<const double2 *>(F.base + k * FD + off);
        const double2 x = __ldg(q);
        r[0] = x.x; r[1] = x.y;
        if (W == 4) { const double2 y = __ldg(q + 1); r[2] = y.x; r[3] = y.y; }
    } else if (FSMEM) {
        // per-problem record staged by TMA: slot rows hold K..Ec (0..45) [+ chat] for the backward sweep and
        // K (0..9), A, B (10..39) [+ c] for the forward sweep; row = off below 46, off - 36 above; 256 B per row
        const uint32_t a = F.sbase + (uint32_t)(off < 46 ? off : off - 36) * 256u;
#pragma unroll
        for (int i = 0; i < W; ++i) r[i] = lds64_ordered(a + 256u * i);
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) r[i] = ADMMB_LD(F.base + ((size_t)k * FD + off + i) * F.ld);
    }
}

// one entry of the packed record (the affine term c)
template <bool FSH, bool FSMEM>
__device__ __forceinline__ double dec_ld1(const FacRef<FSH> &F, int k, int off)
{
    if (FSH) return F.base[k * FD + off];
    if (FSMEM) return lds64_ordered(F.sbase + (uint32_t)(off < 46 ? off : off - 36) * 256u);
    return F.base[((size_t)k * FD + off) * F.ld];
}

// PD = prefetch distance in stages (even): 2 under the 128-register cap, 4 in the uncapped build, where a
// lone warp per sub-partition has nothing else to hide the global-load latency behind.
template <bool FSH, bool FSMEM, bool HAS_C, bool HAS_Q, bool ADAPT, int PD, class Staging = NoStaging, class IO = GlobalIO>
__device__ __forceinline__ void admm_iteration_dec(const IterParams &P, const size_t p, FacRef<FSH> F,
                                                   const int *bdesc, const uint32_t par_sbase, const double rho,
                                                   const double sigma, double (&nr)[5], Staging stg = Staging(),
                                                   const bool act = true, const IO io = IO())
{
    stg.iter_begin(F);
    const int N = P.N;
    const size_t ld = P.ld;
    const double rinv = 1.0 / rho;
    const size_t ldi = io.pitch(P);                   // pitch of the iterate rows (z, u, d)
    const ptrdiff_t ld1 = (ptrdiff_t)ldi, ld2 = 2 * (ptrdiff_t)ldi, ld3 = 3 * (ptrdiff_t)ldi;
    double *const zp = io.z(P, p), *const up = io.u(P, p), *const dp = io.d(P, p);
    const double *qp = HAS_Q ? (P.q_batched ? P.q + p : P.q) : nullptr;
    const size_t qld = P.q_batched ? ld : 1;

    auto load_par = [&](int b, double (&pr)[8]) {
        if (P.par_batched) {
            const double *pp = P.par + p + (size_t)(8 * b) * ld;
#pragma unroll
            for (int q = 0; q < 8; ++q) pr[q] = pp[(size_t)q * ld];
        } else {
            const uint32_t a = par_sbase + (uint32_t)b * 64u;
            const double2 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32), w = lds128(a + 48);
            pr[0] = x.x; pr[1] = x.y; pr[2] = y.x; pr[3] = y.y; pr[4] = z.x; pr[5] = z.y; pr[6] = w.x; pr[7] = w.y;
        }
    };
    auto rt_terminal = [&](int b, double (&t)[3]) {
        const int de = bdesc[b];
        if ((de & 0xff) != BLK_NONE) {
            const size_t r0 = (size_t)(de >> 8) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double uu = IO::ld(up + (r0 + e) * ldi);
                if (ADAPT) uu = uu * sigma;
                double v = IO::ld(zp + (r0 + e) * ldi) - uu;
                if (HAS_Q) v = fma(-qp[(size_t)(3 * b + e) * qld], rinv, v);
                t[e] = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) t[e] = HAS_Q ? -(qp[(size_t)(3 * b + e) * qld] * rinv) : 0.0;
        }
    };

    // ---------------- backward sweep.  g is kept split: gi = (g0,g1,g3,g4) in-plane, gc = (g2,g5) cross-track
    double giA[4], gcA[2], giB[4], gcB[2];
    {
        double t0[3], t1[3];
        rt_terminal(3 * N, t0);
        rt_terminal(3 * N + 1, t1);
        giA[0] = t0[0]; giA[1] = t0[1]; gcA[0] = t0[2]; giA[2] = t1[0]; giA[3] = t1[1]; gcA[1] = t1[2];
    }
    double zb[PD][3], ub[PD][3];                       // slot u holds stage (top - u) of the current group
#pragma unroll
    for (int t = 0; t < PD; ++t)
        if (N - 1 - t >= 0) {
            const double *z0 = zp + (ptrdiff_t)(N - 1 - t) * ld3, *u0 = up + (ptrdiff_t)(N - 1 - t) * ld3;
            zb[t][0] = IO::ld(z0); zb[t][1] = IO::ld(z0 + ld1); zb[t][2] = IO::ld(z0 + ld2);
            ub[t][0] = IO::ld(u0); ub[t][1] = IO::ld(u0 + ld1); ub[t][2] = IO::ld(u0 + ld2);
        }
    const double *zl = zp + (ptrdiff_t)(N - 1 - PD) * ld3, *ul = up + (ptrdiff_t)(N - 1 - PD) * ld3;   // stage k-PD
    double *ds = dp + (ptrdiff_t)(N - 1) * ld3;
    auto bwd_stage = [&](const int k, double (&zc)[3], double (&uc)[3], const double (&gi_in)[4],
                         const double (&gc_in)[2], double (&pi)[4], double (&pc)[2]) {
        stg.bwd_begin(k, F);
        double ra[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            double uu = ADAPT ? uc[e] * sigma : uc[e];
            double v = zc[e] - uu;
            if (HAS_Q) v = fma(-qp[(size_t)(9 * k + 6 + e) * qld], rinv, v);
            ra[e] = v;
        }
        if (k >= PD) {   // refill this slot with the stage PD steps ahead
            zc[0] = IO::ld(zl); zc[1] = IO::ld(zl + ld1); zc[2] = IO::ld(zl + ld2);
            uc[0] = IO::ld(ul); uc[1] = IO::ld(ul + ld1); uc[2] = IO::ld(ul + ld2);
        }
        zl -= ld3; ul -= ld3;
        double gi[4], gc[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) gi[i] = gi_in[i];
        gc[0] = gc_in[0]; gc[1] = gc_in[1];
        if (HAS_C) {
            double c0[4], c1[2];
            dec_ld<FSH, FSMEM, 4>(F, k, D_CHAT, c0);       // chat[0..3]
            dec_ld<FSH, FSMEM, 2>(F, k, D_CHAT + 4, c1);   // chat[4..5]
            gi[0] = gi[0] - c0[0]; gi[1] = gi[1] - c0[1]; gc[0] = gc[0] - c0[2];
            gi[2] = gi[2] - c0[3]; gi[3] = gi[3] - c1[0]; gc[1] = gc[1] - c1[1];
        }
        // rs = -q/rho on the (unsplit) state entries, natural indices 0,1,3,4 | 2,5
        pi[0] = HAS_Q ? -(qp[(size_t)(9 * k + 0) * qld] * rinv) : 0.0;
        pi[1] = HAS_Q ? -(qp[(size_t)(9 * k + 1) * qld] * rinv) : 0.0;
        pi[2] = HAS_Q ? -(qp[(size_t)(9 * k + 3) * qld] * rinv) : 0.0;
        pi[3] = HAS_Q ? -(qp[(size_t)(9 * k + 4) * qld] * rinv) : 0.0;
        pc[0] = HAS_Q ? -(qp[(size_t)(9 * k + 2) * qld] * rinv) : 0.0;
        pc[1] = HAS_Q ? -(qp[(size_t)(9 * k + 5) * qld] * rinv) : 0.0;
        double dj[3];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double h[2], er[4];
            dec_ld<FSH, FSMEM, 2>(F, k, D_HIN + 2 * j, h);
            dec_ld<FSH, FSMEM, 4>(F, k, D_EIN + 4 * j, er);
            double acc = h[0] * ra[0];
            acc = fma(h[1], ra[1], acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc = fma(er[i], gi[i], acc);
            dj[j] = acc;
        }
        {
            double h[2], er[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_HC, h);
            dec_ld<FSH, FSMEM, 2>(F, k, D_EC, er);
            double acc = h[0] * ra[2];
            acc = fma(er[0], gc[0], acc);
            acc = fma(er[1], gc[1], acc);
            dj[2] = acc;
        }
        if (act) { IO::st(ds, dj[0]); IO::st(ds + ld1, dj[1]); IO::st(ds + ld2, dj[2]); }
        ds -= ld3;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double kr[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_KIN + 4 * j, kr);
#pragma unroll
            for (int i = 0; i < 4; ++i) pi[i] = fma(kr[i], ra[j], pi[i]);
        }
        {
            double kr[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_KC, kr);
            pc[0] = fma(kr[0], ra[2], pc[0]);
            pc[1] = fma(kr[1], ra[2], pc[1]);
        }
        // Acl' g: rows l of Acl in ascending natural order 0,1,(2),3,4,(5); in-plane and cross-track
        // accumulators are disjoint, so the two groups can be walked separately without reordering any sum
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            double ar[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_ACLIN + 4 * l, ar);
#pragma unroll
            for (int i = 0; i < 4; ++i) pi[i] = fma(ar[i], gi[l], pi[i]);
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            double ar[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_ACLC + 2 * l, ar);
            pc[0] = fma(ar[0], gc[l], pc[0]);
            pc[1] = fma(ar[1], gc[l], pc[1]);
        }
        stg.stage_end();
    };
    // PD == 2 (register-capped build): one guarded copy of the stage body per slot keeps the code small and the
    // register allocator out of spills.  PD == 4 (uncapped build, latency-bound warps): an unguarded main loop
    // lets the scheduler overlap neighbouring stages; the short remainder is guarded.
    {
        int k = N - 1;
        if (PD > 2) {
            for (; k >= PD - 1; k -= PD) {
#pragma unroll
                for (int t = 0; t < PD; t += 2) {
                    bwd_stage(k - t, zb[t], ub[t], giA, gcA, giB, gcB);
                    bwd_stage(k - t - 1, zb[t + 1], ub[t + 1], giB, gcB, giA, gcA);
                }
            }
        }
        for (; k >= 0; k -= PD) {
#pragma unroll
            for (int t = 0; t < PD; t += 2) {
                if (k - t >= 0) bwd_stage(k - t, zb[t], ub[t], giA, gcA, giB, gcB);
                if (k - t - 1 >= 0) bwd_stage(k - t - 1, zb[t + 1], ub[t + 1], giB, gcB, giA, gcA);
            }
        }
    }

    // ---------------- forward sweep.  s split the same way: si = (s0,s1,s3,s4), sc = (s2,s5)
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    double siA[4], scA[2], siB[4], scB[2];
    siA[0] = io.s0(P, p, 0); siA[1] = io.s0(P, p, 1); scA[0] = io.s0(P, p, 2);
    siA[2] = io.s0(P, p, 3); siA[3] = io.s0(P, p, 4); scA[1] = io.s0(P, p, 5);
    double db[PD][3];
#pragma unroll
    for (int t = 0; t < PD; ++t)
        if (t < N) {
            const double *z0 = zp + (ptrdiff_t)t * ld3, *u0 = up + (ptrdiff_t)t * ld3, *d0 = dp + (ptrdiff_t)t * ld3;
            zb[t][0] = IO::ld(z0); zb[t][1] = IO::ld(z0 + ld1); zb[t][2] = IO::ld(z0 + ld2);
            ub[t][0] = IO::ld(u0); ub[t][1] = IO::ld(u0 + ld1); ub[t][2] = IO::ld(u0 + ld2);
            db[t][0] = IO::ld(d0); db[t][1] = IO::ld(d0 + ld1); db[t][2] = IO::ld(d0 + ld2);
        }
    const double *zf = zp + PD * ld3, *uf = up + PD * ld3, *df = dp + PD * ld3;   // stage k+PD
    double *zw = zp, *uw = up;
    auto fwd_stage = [&](const int k, double (&zc)[3], double (&uc)[3], double (&dc)[3], const double (&si)[4],
                         const double (&sc)[2], double (&ni)[4], double (&nc)[2]) {
        stg.fwd_begin(k, F);
        double a[3], zo[3], uo[3];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double kr[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_KIN + 4 * j, kr);
            double acc = dc[j];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc = fma(kr[i], si[i], acc);
            a[j] = acc;
        }
        {
            double kr[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_KC, kr);
            double acc = dc[2];
            acc = fma(kr[0], sc[0], acc);
            acc = fma(kr[1], sc[1], acc);
            a[2] = acc;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) { zo[j] = zc[j]; uo[j] = ADAPT ? uc[j] * sigma : uc[j]; }
        if (k + PD < N) {
            zc[0] = IO::ld(zf); zc[1] = IO::ld(zf + ld1); zc[2] = IO::ld(zf + ld2);
            uc[0] = IO::ld(uf); uc[1] = IO::ld(uf + ld1); uc[2] = IO::ld(uf + ld2);
            dc[0] = IO::ld(df); dc[1] = IO::ld(df + ld1); dc[2] = IO::ld(df + ld2);
        }
        zf += ld3; uf += ld3; df += ld3;
        {
            const int b = 3 * k + 2;
            double pr[8];
            load_par(b, pr);
            block_update<IO>(bdesc[b] & 0xff, [&](int q) { return pr[q]; }, rinv, P.alpha, P.oma, a, zo, uo, zw, uw, ldi, rr,
                         ss, xx, zz, uu, act);
            zw += ld3; uw += ld3;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double ar[4], br[2];
            dec_ld<FSH, FSMEM, 4>(F, k, D_AIN + 4 * i, ar);
            dec_ld<FSH, FSMEM, 2>(F, k, D_BIN + 2 * i, br);
            double acc = ar[0] * si[0];
            acc = fma(ar[1], si[1], acc);
            acc = fma(ar[2], si[2], acc);
            acc = fma(ar[3], si[3], acc);
            acc = fma(br[0], a[0], acc);
            acc = fma(br[1], a[1], acc);
            if (HAS_C) acc = acc + dec_ld1<FSH, FSMEM>(F, k, D_C + (i < 2 ? i : i + 1));
            ni[i] = acc;
        }
        {
            double bc[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_BC, bc);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double ar[2];
                dec_ld<FSH, FSMEM, 2>(F, k, D_AC + 2 * i, ar);
                double acc = ar[0] * sc[0];
                acc = fma(ar[1], sc[1], acc);
                acc = fma(bc[i], a[2], acc);
                if (HAS_C) acc = acc + dec_ld1<FSH, FSMEM>(F, k, D_C + (i == 0 ? 2 : 5));
                nc[i] = acc;
            }
        }
        stg.stage_end();
    };
    const bool s_in_A = (N & 1) == 0;                   // every stage flips the s ping-pong
    {
        int k = 0;
        if (PD > 2) {
            for (; k + PD <= N; k += PD) {
#pragma unroll
                for (int t = 0; t < PD; t += 2) {
                    fwd_stage(k + t, zb[t], ub[t], db[t], siA, scA, siB, scB);
                    fwd_stage(k + t + 1, zb[t + 1], ub[t + 1], db[t + 1], siB, scB, siA, scA);
                }
            }
        }
        for (; k < N; k += PD) {
#pragma unroll
            for (int t = 0; t < PD; t += 2) {
                if (k + t < N) fwd_stage(k + t, zb[t], ub[t], db[t], siA, scA, siB, scB);
                if (k + t + 1 < N) fwd_stage(k + t + 1, zb[t + 1], ub[t + 1], db[t + 1], siB, scB, siA, scA);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int b = 3 * N + t;
        const int de = bdesc[b];
        if ((de & 0xff) == BLK_NONE) continue;
        const size_t r0 = (size_t)(de >> 8) * 3;
        const double *si = s_in_A ? siA : siB, *sc = s_in_A ? scA : scB;
        const double xb[3] = {si[2 * t], si[2 * t + 1], sc[t]};       // (s0,s1,s2) or (s3,s4,s5)
        double zo[3], uo[3], pr[8];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            zo[e] = IO::ld(zp + (r0 + e) * ldi);
            double uv = IO::ld(up + (r0 + e) * ldi);
            uo[e] = ADAPT ? uv * sigma : uv;
        }
        load_par(b, pr);
        block_update<IO>(de & 0xff, [&](int q) { return pr[q]; }, rinv, P.alpha, P.oma, xb, zo, uo, zp + r0 * ldi,
                     up + r0 * ldi, ldi, rr, ss, xx, zz, uu, act);
    }
    nr[0] = rr; nr[1] = ss; nr[2] = xx; nr[3] = zz; nr[4] = uu;
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous global array into shared memory,
// completion signalled on an mbarrier: one instruction stages the whole 62 KB factor table.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Rows a5 + a6: persistent multi-iteration launch.  Each thread runs up to `chunk` iterations of
// its problem, evaluates the stopping test on device, leaves as soon as its problem is done
// (per-problem early exit) and applies the residual-balancing rho update, refactorising only its
// own problem when the factor depends on rho.
// dynamic smem: [16 B mbarrier][FSH && FSMEM ? FS*N doubles : 0][par shared ? 8*nb doubles : 0][nb ints]
// LOWOCC: variant compiled without the 128-register cap, used when the active set is small enough
// that occupancy does not matter (no spills of the prefetch registers).
// ------------------------------------------------------------------------------------------------
template <bool FSH, bool FSMEM, bool HAS_C, bool HAS_Q, bool ADAPT, int MODE, bool LOWOCC>
__global__ void __launch_bounds__(256, LOWOCC ? 1 : 2) k_admm_iterate(const __grid_constant__ IterParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *facS = reinterpret_cast<double *>(smem_raw + 16);
    constexpr int REC = (MODE == 2) ? FD : FS;                  // doubles per staged factor record
    const double *fac_src = (MODE == 2) ? P.fac_dec : P.fac;
    double *parS = facS + ((FSH && FSMEM) ? (size_t)REC * P.N : 0);
    int *bdS = reinterpret_cast<int *>(parS + (P.par_batched ? 0 : 8 * P.nb));
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t fac_sbase = (uint32_t)__cvta_generic_to_shared(facS);
    const uint32_t par_sbase = (uint32_t)__cvta_generic_to_shared(parS);
    const uint32_t bd_bytes = (uint32_t)(((P.nb + 3) / 4) * 16);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        uint32_t bytes = bd_bytes;
        if (FSH && FSMEM) bytes += (uint32_t)(REC * P.N * 8);
        if (!P.par_batched) bytes += (uint32_t)(8 * P.nb * 8);
        mbar_expect_tx(mbar, bytes);
        if (FSH && FSMEM) bulk_g2s(fac_sbase, fac_src, (uint32_t)(REC * P.N * 8), mbar);
        if (!P.par_batched) bulk_g2s(par_sbase, P.par, (uint32_t)(8 * P.nb * 8), mbar);
        bulk_g2s((uint32_t)__cvta_generic_to_shared(bdS), P.bdesc, bd_bytes, mbar);
    }
    __syncthreads();
    mbar_wait(mbar, 0);

    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < P.n_active && P.status[t] == ST_RUNNING;
    const unsigned wmask = __ballot_sync(0xffffffffu, live);   // the lanes of this warp that run the loop
    if (!live) return;
    const size_t p = (size_t)t;       // column of the (densely repacked) working set: always aligned

    FacRef<FSH> F;
    F.base = FSH ? (FSMEM ? facS : fac_src) : fac_src + p;
    F.ld = P.ld;
    F.sbase = fac_sbase;

    double rho = P.rho[p];
    double sigma = ADAPT ? P.usc[p] : 1.0;
    int it = P.iters[p];
    int st = ST_RUNNING;
    double r_norm = 0.0, s_norm = 0.0, eps_pri = 0.0, eps_dual = 0.0;
    // A warp in which some lane has left the loop runs the remaining iterations ~25 % slower (measured: 1.62 ms per 50
    // iterations with full warps, 2.02-2.09 ms as soon as one lane of one warp is idle; presumably because its stores
    // no longer cover whole 32-byte sectors -- DESIGN 4.1c) -- and on a narrow working set the slowest warp is the
    // launch time.  So a lane whose problem finishes does not leave: it copies its final z, u, d to the
    // home columns (snapshot), then keeps iterating on its working column (whose contents no longer matter) until
    // every lane of the warp is done or the launch ends.  Needs a working set that is a copy (P.z_home != nullptr).
    const bool zombies = P.z_home != nullptr;
    bool done = false;
    double sigma_fin = 1.0;
    for (int cnt = 0; cnt < P.chunk; ++cnt) {
        if (!done) ++it;
        double nr[5];
        if (MODE == 2) admm_iteration_dec<FSH, FSMEM, HAS_C, HAS_Q, ADAPT, (LOWOCC ? 4 : 2)>(P, p, F, bdS, par_sbase, rho, sigma, nr);
        else if (MODE == 1) admm_iteration_fast<FSH, FSMEM, HAS_C, HAS_Q, ADAPT>(P, p, F, bdS, par_sbase, rho, sigma, nr);
        else admm_iteration<FSH, HAS_C, HAS_Q, ADAPT>(P, p, F, bdS, parS, rho, sigma, nr);
        sigma = 1.0;
        if (!done) {
            r_norm = sqrt(nr[0]);
            s_norm = rho * sqrt(nr[1]);
            const double nx = sqrt(nr[2]), nz = sqrt(nr[3]);
            eps_pri = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
            eps_dual = fma(P.reltol, rho * sqrt(nr[4]), P.sqrtn_abs);
            if (P.hist) {
                const size_t h = (size_t)(it - 1) * P.hist_ld + (P.orig ? (size_t)P.orig[p] : p);   // home column
                P.hist[h] = r_norm;
                P.hist[h + P.hist_stride] = s_norm;
                P.hist[h + 2 * P.hist_stride] = eps_pri;
                P.hist[h + 3 * P.hist_stride] = eps_dual;
                P.hist[h + 4 * P.hist_stride] = rho;
            }
            if (!(isfinite(r_norm) && isfinite(s_norm))) st = ST_NAN;
            else if (r_norm < eps_pri && s_norm < eps_dual) st = ST_CONVERGED;
            else {
                if (ADAPT && (it % P.every) == 0 && it < P.max_iter && (P.until <= 0 || it <= P.until)) {
                    bool ch = false;
                    if (r_norm > P.mu * s_norm) {
                        if (!(rho * P.tau > RHO_MAX)) { rho = rho * P.tau; sigma = P.inv_tau; ch = true; }
                    } else if (s_norm > P.mu * r_norm) {
                        if (!(rho * P.inv_tau < RHO_MIN)) { rho = rho * P.inv_tau; sigma = P.tau; ch = true; }
                    }
                    if (ch && !FSH && P.has_P) {
                        const size_t off = P.raw_batched ? p : 0, ldr = P.raw_batched ? P.ld : 1;
                        int bad = riccati_factor_dev(P.N, P.rawA + off, P.rawB + off, P.rawc ? P.rawc + off : nullptr,
                                                     P.rawQ ? P.rawQ + off : nullptr, P.rawR ? P.rawR + off : nullptr,
                                                     ldr, rho, bdS, P.fac_rw + p, P.ld);
                        if ((int)p < P.n_real) atomicAdd(P.refac_count, 1ULL);
                        if (bad) st = ST_NAN;
                        else if (MODE == 2) pack_decoupled_dev(P.N, P.fac_rw + p, P.ld, P.fac_dec_rw + p, P.ld);
                    }
                }
                if (st == ST_RUNNING && it >= P.max_iter) st = ST_MAX_ITER;
            }
            if (st != ST_RUNNING) {
                done = true;
                sigma_fin = sigma;
                if (!zombies) break;
                const size_t h = (size_t)P.orig[p];
                for (int r = 0; r < P.rows_zu; ++r) {
                    P.z_home[(size_t)r * P.home_ld + h] = P.z[(size_t)r * P.ld + p];
                    P.u_home[(size_t)r * P.home_ld + h] = P.u[(size_t)r * P.ld + p];
                }
                for (int r = 0; r < 3 * P.N; ++r) P.d_home[(size_t)r * P.home_ld + h] = P.d[(size_t)r * P.ld + p];
                P.snap[p] = 1;
                sigma = 1.0;
            }
        }
        if (zombies && __all_sync(wmask, done)) break;
    }
    if (done) sigma = sigma_fin;
    P.iters[p] = it;
    P.rho[p] = rho;
    if (ADAPT) P.usc[p] = sigma;
    P.status[p] = st;
    P.fin[p] = r_norm;
    P.fin[p + P.ld] = s_norm;
    P.fin[p + 2 * P.ld] = eps_pri;
    P.fin[p + 3 * P.ld] = eps_dual;
}

#ifndef ADMMB_ITERATE_ONLY
// ------------------------------------------------------------------------------------------------
// Physical compaction between launches.  A compacted INDEX list leaves warps reading 32 problems that
// are no longer 256-byte aligned; measured on B200 a shift by ONE problem costs 35 % and a fragmented
// list up to 4x.  So the still-running problems are physically repacked into a dense, aligned prefix
// (k_split -> k_gather_cols) and the finished ones are retired to their home columns (k_scatter_cols).
// ------------------------------------------------------------------------------------------------
__global__ void k_split(const int *status, int n, int *keep, int *fin, int *counts)
{
    __shared__ int wk[32], wf[32];
    __shared__ int base_k, base_f;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool valid = t < n;
    const bool k = valid && status[t] == ST_RUNNING, f = valid && !k;
    const unsigned mk = __ballot_sync(0xffffffffu, k), mf = __ballot_sync(0xffffffffu, f);
    if (lane == 0) { wk[wid] = __popc(mk); wf[wid] = __popc(mf); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tk = 0, tf = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            int c = wk[w]; wk[w] = tk; tk += c;
            c = wf[w]; wf[w] = tf; tf += c;
        }
        base_k = tk ? atomicAdd(counts, tk) : 0;
        base_f = tf ? atomicAdd(counts + 1, tf) : 0;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1;
    if (k) keep[base_k + wk[wid] + __popc(mk & below)] = t;
    if (f) fin[base_f + wf[wid] + __popc(mf & below)] = t;
}

// out[r][t] = in[r][src[t]], t < n  (rows on blockIdx.y)
template <typename T>
__global__ void k_gather_cols(const T *in, size_t ld_in, int rows, const int *src, int n, T *out, size_t ld_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t c = (size_t)src[t];
    for (int r = blockIdx.y; r < rows; r += gridDim.y) out[(size_t)r * ld_out + t] = in[(size_t)r * ld_in + c];
}

// home[r][orig[c]] = in[r][c] for the finished columns c = fin[t]
template <typename T>
__global__ void k_scatter_cols(const T *in, size_t ld_in, int rows, const int *fin, int n, const int *orig, T *home,
                               size_t ld_home, const int *skip = nullptr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t c = (size_t)fin[t], h = (size_t)orig[c];
    if (skip && skip[c]) return;      // already copied home by the kernel at the moment the problem finished
    for (int r = blockIdx.y; r < rows; r += gridDim.y) home[(size_t)r * ld_home + h] = in[(size_t)r * ld_in + c];
}

// orig_new[t] = orig_old ? orig_old[src[t]] : src[t]
__global__ void k_compose_orig(const int *orig_old, const int *src, int n, int *orig_new)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) orig_new[t] = orig_old ? orig_old[src[t]] : src[t];
}

__global__ void k_add_int(int *a, int v) { *a += v; }
__global__ void k_set_int(int *a, int v) { *a = v; }

// keep[n .. n_pad) = keep[n-1]: the working set is padded to whole warps with copies of a running problem
__global__ void k_pad_list(int *keep, int n, int n_pad)
{
    const int t = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_pad) keep[t] = keep[n - 1];
}

__global__ void k_iota(int *a, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = t;
}

// ------------------------------------------------------------------------------------------------
// state reset before a run: iterates from the warm start (or zero), counters, rho
// ------------------------------------------------------------------------------------------------
__global__ void k_reset(int64_t batch, size_t ld, int rows_zu, double *z, double *u, const double *z0,
                        const double *u0, double *rho, const double *rho0, double rho_shared, double *usc,
                        int *iters, int *status, const int *fac_status, const int *active = nullptr)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    // SCP passes (scp.cuh): a problem whose trajectory has converged keeps the state of its last solve -- iterates, iteration
    // count and a status other than RUNNING, so that no kernel touches it again
    if (active && !active[p]) return;
    for (int r = 0; r < rows_zu; ++r) {
        z[(size_t)r * ld + p] = z0 ? z0[(size_t)r * ld + p] : 0.0;
        u[(size_t)r * ld + p] = u0 ? u0[(size_t)r * ld + p] : 0.0;
    }
    rho[p] = rho0 ? rho0[p] : rho_shared;
    usc[p] = 1.0;
    iters[p] = 0;
    status[p] = (fac_status && fac_status[p] == ST_NAN) ? ST_NAN : ST_RUNNING;
}

// ------------------------------------------------------------------------------------------------
// final outputs: x is re-propagated from the stored controls a_k (same operations as the forward
// sweep => same bits), z/u are expanded from the compact split rows (BLK_NONE: z = x, u = 0).
// Writes [n][ld] interleaved arrays; k_transpose_out turns them into [n x batch] column-major.
// ------------------------------------------------------------------------------------------------
template <bool FSH, bool HAS_C>
__global__ void k_output(int N, int64_t batch, size_t ld, const double *fac, const double *s0,
                         const double *d, const double *z, const double *u, const double *usc,
                         const int *bdesc, const int *iters, double *xo, double *zo, double *uo)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    FacRef<FSH> F;
    F.base = FSH ? fac : fac + p;
    F.ld = ld;
    const double sigma = usc ? usc[p] : 1.0;
    const bool ran = iters[p] > 0;
    auto emit = [&](int b, double x0, double x1, double x2) {
        const int de = bdesc[b];
        const double xb[3] = {x0, x1, x2};
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t o = (size_t)(3 * b + e) * ld + p;
            double zv = xb[e], uv = 0.0;
            if ((de & 0xff) != BLK_NONE) {
                const size_t r = ((size_t)(de >> 8) * 3 + e) * ld + p;
                zv = z[r];
                uv = u[r] * sigma;
            }
            if (xo) xo[o] = ran ? xb[e] : 0.0;
            if (zo) zo[o] = ran ? zv : (((de & 0xff) != BLK_NONE) ? zv : 0.0);
            if (uo) uo[o] = uv;
        }
    };
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = s0[p + (size_t)i * ld];
    for (int k = 0; k < N; ++k) {
        double a[3], sn[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {      // a_k = d_k + K_k s_k, the forward sweep's operations
            double acc = d[p + (size_t)(3 * k + j) * ld];
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_K + 6 * j + i), s[i], acc);
            a[j] = acc;
        }
        emit(3 * k, s[0], s[1], s[2]);
        emit(3 * k + 1, s[3], s[4], s[5]);
        emit(3 * k + 2, a[0], a[1], a[2]);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = F(k, F_A + 6 * i + 0) * s[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(F(k, F_A + 6 * i + l), s[l], acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_B + B_LD * i + j), a[j], acc);
            if (HAS_C) acc = acc + F(k, F_C + i);
            sn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = sn[i];
    }
    emit(3 * N, s[0], s[1], s[2]);
    emit(3 * N + 1, s[3], s[4], s[5]);
}

// ------------------------------------------------------------------------------------------------
// layout changes.  in: [batch][R] (each problem's R values contiguous, the MATLAB layout)
//                  out: [R'][ld]   (row rowmap[r], or r when rowmap == nullptr; -1 rows dropped)
// ------------------------------------------------------------------------------------------------
__global__ void k_transpose_in(const double *in, int64_t batch, int R, double *out, size_t ld,
                               const int *rowmap)
{
    __shared__ double tile[32][33];
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t p = p0 + i;
        int r = r0 + threadIdx.x;
        if (p < batch && r < R) tile[i][threadIdx.x] = in[(size_t)p * R + r];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i;
        int64_t p = p0 + threadIdx.x;
        if (p < batch && r < R) {
            int ro = rowmap ? rowmap[r] : r;
            if (ro >= 0) out[(size_t)ro * ld + p] = tile[threadIdx.x][i];
        }
    }
}

// in: [R][ld] interleaved  ->  out: [batch][R]
template <typename T>
__global__ void k_transpose_out(const T *in, size_t ld, int R, int64_t batch, T *out)
{
    __shared__ T tile[32][33];
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i;
        int64_t p = p0 + threadIdx.x;
        if (p < batch && r < R) tile[i][threadIdx.x] = in[(size_t)r * ld + p];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t p = p0 + i;
        int r = r0 + threadIdx.x;
        if (p < batch && r < R) out[(size_t)p * R + r] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------
// standalone kernels (unit entry points and the dense path)
// ------------------------------------------------------------------------------------------------
// Row a2 alone: x = Riccati x-update of rt, all arrays [rows][ld]; shared factor in global memory.
template <bool HAS_C>
__global__ void k_xupdate_riccati(int N, int64_t batch, size_t ld, const double *fac, const double *s0,
                                  const double *rt, double *d, double *x)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    FacRef<true> F;
    F.base = fac;
    F.ld = ld;
    double g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = rt[(size_t)(9 * N + i) * ld + p];
    for (int k = N - 1; k >= 0; --k) {
        double rs[6], ra[3], pn[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) rs[i] = rt[(size_t)(9 * k + i) * ld + p];
#pragma unroll
        for (int j = 0; j < 3; ++j) ra[j] = rt[(size_t)(9 * k + 6 + j) * ld + p];
        if (HAS_C) {
#pragma unroll
            for (int i = 0; i < 6; ++i) g[i] = g[i] - F(k, F_CHAT + i);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = F(k, F_HINV + HINV_LD * j + 0) * ra[0];
            acc = fma(F(k, F_HINV + HINV_LD * j + 1), ra[1], acc);
            acc = fma(F(k, F_HINV + HINV_LD * j + 2), ra[2], acc);
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_E + 6 * j + i), g[i], acc);
            d[(size_t)(3 * k + j) * ld + p] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = rs[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_K + 6 * j + i), ra[j], acc);
#pragma unroll
            for (int l = 0; l < 6; ++l) acc = fma(F(k, F_ACL + 6 * l + i), g[l], acc);
            pn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = pn[i];
    }
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = s0[p + (size_t)i * ld];
    for (int k = 0; k < N; ++k) {
        double a[3], sn[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = d[(size_t)(3 * k + j) * ld + p];
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_K + 6 * j + i), s[i], acc);
            a[j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) x[(size_t)(9 * k + i) * ld + p] = s[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) x[(size_t)(9 * k + 6 + j) * ld + p] = a[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = F(k, F_A + 6 * i + 0) * s[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(F(k, F_A + 6 * i + l), s[l], acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_B + B_LD * i + j), a[j], acc);
            if (HAS_C) acc = acc + F(k, F_C + i);
            sn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = sn[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[(size_t)(9 * N + i) * ld + p] = s[i];
}

// Exact FP64 Riccati solves of the condensed TF32 path (same arithmetic as k_xupdate_riccati), one thread per problem.
//   MODE 0 (retire, dense layout): for the finished working-set columns c = fin[t], rt = (z - u) - (inc_hi + inc_lo) on
//          the split rows (the right-hand side the problem's last iteration used), x written to the home column orig[c].
//   MODE 1 (refresh): for every running column, rt = z - u (the right-hand side of the NEXT iteration), x written to the
//          split rows of the accumulated x_R (out = xacc, pitch ld_out); the caller then zeroes the increment buffers.
//   MODE 2 (retire, Riccati layout): as MODE 0 but only the backward sweep; d is left in the working set (out = the
//          travelling d array, column c) for k_output.
// zu_compact: z / u rows are the compact split rows (Riccati layout) instead of the full rows.  d: scratch [3N][ld_d],
// one column per t (MODE 0 / 1).
template <bool HAS_C, int MODE>
__global__ void k_tf32_final_x(int N, const double *__restrict__ fac, const int *__restrict__ bdesc, const double *s0,
                               const double *z, const double *u, int zu_compact, const float *inc_hi, const float *inc_lo,
                               size_t ld_in, const int *fin, int n_fin, const int *orig, double *d, size_t ld_d, double *out,
                               size_t ld_out, const int *status)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_fin) return;
    if (MODE == 1 && status[t] != ST_RUNNING) return;     // finished columns keep the x_R of their last iteration
    const size_t c = MODE == 1 ? (size_t)t : (size_t)fin[t];
    const size_t h = MODE == 0 ? (orig ? (size_t)orig[c] : c) : c;
    const size_t dcol = MODE == 2 ? c : (size_t)t;
    FacRef<true> F;
    F.base = fac;
    F.ld = 0;
    auto put = [&](int row, double v) {
        if (MODE == 0) { out[(size_t)row * ld_out + h] = v; return; }
        const int bd = bdesc[row / 3];
        if ((bd & 0xff) != BLK_NONE) out[(size_t)(3 * (bd >> 8) + row % 3) * ld_out + h] = v;
    };
    auto rt = [&](int row) -> double {
        const int bd = bdesc[row / 3];
        if ((bd & 0xff) == BLK_NONE) return 0.0;
        const size_t oc = (size_t)(3 * (bd >> 8) + row % 3) * ld_in + c;
        const size_t o = zu_compact ? oc : (size_t)row * ld_in + c;
        if (MODE == 1) return z[o] - u[o];
        const double inc = (double)inc_hi[oc] + (inc_lo ? (double)inc_lo[oc] : 0.0);
        return (z[o] - u[o]) - inc;
    };
    double g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = rt(9 * N + i);
    for (int k = N - 1; k >= 0; --k) {
        double rs[6], ra[3], pn[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) rs[i] = rt(9 * k + i);
#pragma unroll
        for (int j = 0; j < 3; ++j) ra[j] = rt(9 * k + 6 + j);
        if (HAS_C) {
#pragma unroll
            for (int i = 0; i < 6; ++i) g[i] = g[i] - F(k, F_CHAT + i);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = F(k, F_HINV + HINV_LD * j + 0) * ra[0];
            acc = fma(F(k, F_HINV + HINV_LD * j + 1), ra[1], acc);
            acc = fma(F(k, F_HINV + HINV_LD * j + 2), ra[2], acc);
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_E + 6 * j + i), g[i], acc);
            d[(size_t)(3 * k + j) * ld_d + dcol] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = rs[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_K + 6 * j + i), ra[j], acc);
#pragma unroll
            for (int l = 0; l < 6; ++l) acc = fma(F(k, F_ACL + 6 * l + i), g[l], acc);
            pn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) g[i] = pn[i];
    }
    if constexpr (MODE != 2) {   // MODE 2 keeps d only (k_output runs the forward sweep at download)
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = s0[(size_t)i * ld_in + c];
    for (int k = 0; k < N; ++k) {
        double a[3], sn[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = d[(size_t)(3 * k + j) * ld_d + dcol];
#pragma unroll
            for (int i = 0; i < 6; ++i) acc = fma(F(k, F_K + 6 * j + i), s[i], acc);
            a[j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) put(9 * k + i, s[i]);
#pragma unroll
        for (int j = 0; j < 3; ++j) put(9 * k + 6 + j, a[j]);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = F(k, F_A + 6 * i + 0) * s[0];
#pragma unroll
            for (int l = 1; l < 6; ++l) acc = fma(F(k, F_A + 6 * i + l), s[l], acc);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fma(F(k, F_B + B_LD * i + j), a[j], acc);
            if (HAS_C) acc = acc + F(k, F_C + i);
            sn[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = sn[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) put(9 * N + i, s[i]);
    }
}

// Rows a3 + a4 alone (the streaming prox / dual / residual kernel, "C4"): reads x, z, u once,
// writes z, u once: 40 bytes per split entry.  All arrays are full-width [n][ld]; BLK_NONE rows are
// skipped.  Optionally emits the next right-hand side rt = w*(z-u) - q/rho for the dense x-update
// and, with SOLVE, evaluates the stopping test on device (per-problem early exit on the dense path).
struct DenseStep {
    int it;                    // iteration number of this step
    int max_iter;
    double reltol, sqrtn_abs;
    const double *rho;         // [ld]
    int *iters, *status;       // [ld]
    double *fin;               // [4][ld]
    int *running;              // incremented once per problem still running after this step
    const int *it_base;        // optional (CUDA-graph replay): iteration number = it + *it_base
};

// F32: x comes from the TF32 tensor-core GEMM as fp32 and the next right-hand side is written as fp32
// (hi part = TF32-representable head, optional lo part for the 3xTF32 split); a problem that finishes
// copies its x column to the FP64 output buffer x_final.
template <bool SOLVE, bool F32>
__global__ void k_prox_dual_residuals(int nb, int64_t batch, size_t ld, const int *bdesc, const double *par,
                                      int par_batched, const double *rinv_arr, double rinv_shared,
                                      double alpha, const void *xin, double *z, double *u, double *norms,
                                      void *rt_out, float *rt_lo, double *x_final, const double *q, int q_batched,
                                      const DenseStep ds)
{
    const double *x = F32 ? nullptr : (const double *)xin;
    const float *x32 = F32 ? (const float *)xin : nullptr;
    double *rt_next = F32 ? nullptr : (double *)rt_out;
    float *rt_hi = F32 ? (float *)rt_out : nullptr;
    auto put_rt = [&](size_t o, double t) {
        if (F32) {
            const float f = (float)t;
            const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
            rt_hi[o] = h;
            if (rt_lo) rt_lo[o] = (float)(t - (double)h);
        } else {
            rt_next[o] = t;
        }
    };
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    if (SOLVE && ds.status[p] != ST_RUNNING) return;
    const double rho = SOLVE ? ds.rho[p] : 0.0;
    const double rinv = SOLVE ? 1.0 / rho : (rinv_arr ? rinv_arr[p] : rinv_shared);
    const double oma = 1.0 - alpha;
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    for (int b = 0; b < nb; ++b) {
        const int type = bdesc[b] & 0xff;
        if (type == BLK_NONE) {
            // unsplit rows of the right-hand side are -q/rho: constant (zero) without a linear cost, written once by
            // k_dense_rt_init -- rewriting them every iteration was 25 % of this kernel's traffic
            if (rt_out && q)
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const size_t o = (size_t)(3 * b + e) * ld + p;
                    put_rt(o, -(q[q_batched ? o : (size_t)(3 * b + e)] * rinv));
                }
            continue;
        }
        double xb[3], zo[3], v[3], zn[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t o = (size_t)(3 * b + e) * ld + p;
            xb[e] = F32 ? (double)x32[o] : ld_stream(x + o);
            zo[e] = ld_stream(z + o);
            double uo = ld_stream(u + o);
            double xh = fma(alpha, xb[e], oma * zo[e]);
            v[e] = xh + uo;
        }
        if (par_batched) {
            const double *pp = par + p + (size_t)(8 * b) * ld;
            prox_block_dev(type, [&](int s) { return pp[(size_t)s * ld]; }, rinv, v, zn);
        } else {
            const double *pp = par + 8 * b;
            prox_block_dev(type, [&](int s) { return __ldg(pp + s); }, rinv, v, zn);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t o = (size_t)(3 * b + e) * ld + p;
            double un = v[e] - zn[e];
            double dr = xb[e] - zn[e];
            double dz = zn[e] - zo[e];
            rr = fma(dr, dr, rr);
            ss = fma(dz, dz, ss);
            xx = fma(xb[e], xb[e], xx);
            zz = fma(zn[e], zn[e], zz);
            uu = fma(un, un, uu);
            st_stream(z + o, zn[e]);
            st_stream(u + o, un);
            if (rt_out) {
                double t = zn[e] - un;
                if (q) t = fma(-q[q_batched ? o : (size_t)(3 * b + e)], rinv, t);
                put_rt(o, t);
            }
        }
    }
    if (norms) {
        norms[p] = rr;
        norms[ld + p] = ss;
        norms[2 * ld + p] = xx;
        norms[3 * ld + p] = zz;
        norms[4 * ld + p] = uu;
    }
    if (SOLVE) {
        const double r_norm = sqrt(rr), s_norm = rho * sqrt(ss);
        const double nx = sqrt(xx), nz = sqrt(zz);
        const double eps_pri = fma(ds.reltol, nx > nz ? nx : nz, ds.sqrtn_abs);
        const double eps_dual = fma(ds.reltol, rho * sqrt(uu), ds.sqrtn_abs);
        int st = ST_RUNNING;
        if (!(isfinite(r_norm) && isfinite(s_norm))) st = ST_NAN;
        else if (r_norm < eps_pri && s_norm < eps_dual) st = ST_CONVERGED;
        else if (ds.it >= ds.max_iter) st = ST_MAX_ITER;
        ds.iters[p] = ds.it;
        ds.status[p] = st;
        ds.fin[p] = r_norm;
        ds.fin[p + ld] = s_norm;
        ds.fin[p + 2 * ld] = eps_pri;
        ds.fin[p + 3 * ld] = eps_dual;
        if (st == ST_RUNNING) atomicAdd(ds.running, 1);
        else if (F32 && x_final)
            for (int r = 0; r < 3 * nb; ++r) x_final[(size_t)r * ld + p] = (double)x32[(size_t)r * ld + p];
    }
}

// Condensed incremental TF32 dense path: C3 + C4 + stop test over the split blocks only, several threads per
// problem.  blockDim = (32 problems, CH block-chunks): thread (tx, ty) handles split blocks ty, ty + CH, ... of
// problem blockIdx.x * 32 + tx, so a warp still touches 32 consecutive problems of one row (coalesced) while the
// CH warps of a CTA spread the per-problem serial work; the partial norms meet in shared memory and are summed
// in ascending ty order (deterministic).  dx (the tensor-core product M_RR * increment), xacc and the increment
// use compact rows (3j+e for split block j); z, u and par the full rows / block numbers.
//   x = xacc + dx (FP64 accumulate),   increment for the next GEMM = (z+ - z) - (u+ - u)
template <int CH, int MINW>   // MINW: resident warps per SM the register allocation must allow
__global__ void __launch_bounds__(32 * CH, MINW / CH)
k_prox_cond_tf32(int nsb, const int *__restrict__ sblk, const int *__restrict__ bdesc, int64_t batch, size_t ld,
                 const double *par, int par_batched, double alpha, const float *__restrict__ dx32, double *xacc,
                 double *z, double *u, int zu_compact, float *rt_hi, float *rt_lo, const DenseStep ds)
{
    __shared__ double red[CH][5][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t p = (int64_t)blockIdx.x * 32 + tx;
    const bool active = p < batch && ds.status[p] == ST_RUNNING;
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    double rho = 1.0;
    if (active) {
        rho = ds.rho[p];
        const double rinv = 1.0 / rho;
        const double oma = 1.0 - alpha;
        for (int j = ty; j < nsb; j += CH) {
            const int b = sblk[j];
            const int type = bdesc[b] & 0xff;
            double xb[3], zo[3], uo[3], v[3], zn[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const size_t oc = (size_t)(3 * j + e) * ld + p, o = zu_compact ? oc : (size_t)(3 * b + e) * ld + p;
                xb[e] = ld_stream(xacc + oc) + (double)__ldcs(dx32 + oc);
                st_stream(xacc + oc, xb[e]);
                zo[e] = ld_stream(z + o);
                uo[e] = ld_stream(u + o);
                const double xh = fma(alpha, xb[e], oma * zo[e]);
                v[e] = xh + uo[e];
            }
            if (par_batched) {
                const double *pp = par + p + (size_t)(8 * b) * ld;
                prox_block_dev(type, [&](int q) { return pp[(size_t)q * ld]; }, rinv, v, zn);
            } else {
                const double *pp = par + 8 * b;
                prox_block_dev(type, [&](int q) { return __ldg(pp + q); }, rinv, v, zn);
            }
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const size_t oc = (size_t)(3 * j + e) * ld + p, o = zu_compact ? oc : (size_t)(3 * b + e) * ld + p;
                const double un = v[e] - zn[e];
                const double dr = xb[e] - zn[e];
                const double dz = zn[e] - zo[e];
                rr = fma(dr, dr, rr);
                ss = fma(dz, dz, ss);
                xx = fma(xb[e], xb[e], xx);
                zz = fma(zn[e], zn[e], zz);
                uu = fma(un, un, uu);
                st_stream(z + o, zn[e]);
                st_stream(u + o, un);
                const double t = dz - (un - uo[e]);     // increment of z - u: differences of nearby numbers first
                const float f = (float)t;
                const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
                rt_hi[oc] = h;
                if (rt_lo) rt_lo[oc] = (float)(t - (double)h);
            }
        }
    }
    red[ty][0][tx] = rr; red[ty][1][tx] = ss; red[ty][2][tx] = xx; red[ty][3][tx] = zz; red[ty][4][tx] = uu;
    __syncthreads();
    if (ty != 0 || !active) return;
#pragma unroll
    for (int c = 1; c < CH; ++c) {
        rr += red[c][0][tx]; ss += red[c][1][tx]; xx += red[c][2][tx]; zz += red[c][3][tx]; uu += red[c][4][tx];
    }
    const int it_now = ds.it + (ds.it_base ? *ds.it_base : 0);
    const double r_norm = sqrt(rr), s_norm = rho * sqrt(ss);
    const double nx = sqrt(xx), nz = sqrt(zz);
    const double eps_pri = fma(ds.reltol, nx > nz ? nx : nz, ds.sqrtn_abs);
    const double eps_dual = fma(ds.reltol, rho * sqrt(uu), ds.sqrtn_abs);
    int st = ST_RUNNING;
    if (!(isfinite(r_norm) && isfinite(s_norm))) st = ST_NAN;
    else if (r_norm < eps_pri && s_norm < eps_dual) st = ST_CONVERGED;
    else if (it_now >= ds.max_iter) st = ST_MAX_ITER;
    ds.iters[p] = it_now;
    ds.status[p] = st;
    ds.fin[p] = r_norm;
    ds.fin[p + ld] = s_norm;
    ds.fin[p + 2 * ld] = eps_pri;
    ds.fin[p + 3 * ld] = eps_dual;
    }

// first right-hand side of the dense path: rt = w*(z - u) - q/rho over full-width rows
__global__ void k_dense_rt_init(int nb, int64_t batch, size_t ld, const int *bdesc, const double *z,
                                const double *u, const double *rho, const double *q, int q_batched, double *rt)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const double rinv = 1.0 / rho[p];
    for (int b = 0; b < nb; ++b) {
        const bool split = (bdesc[b] & 0xff) != BLK_NONE;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t o = (size_t)(3 * b + e) * ld + p;
            double t;
            if (split) {
                t = z[o] - u[o];
                if (q) t = fma(-q[q_batched ? o : (size_t)(3 * b + e)], rinv, t);
            } else {
                t = q ? -(q[q_batched ? o : (size_t)(3 * b + e)] * rinv) : 0.0;
            }
            rt[o] = t;
        }
    }
}

// dense-path outputs: z = x and u = 0 on BLK_NONE rows
__global__ void k_dense_output(int nb, int64_t batch, size_t ld, const int *bdesc, const double *x,
                               const double *z, const double *u, double *zo, double *uo)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    for (int b = 0; b < nb; ++b) {
        const bool split = (bdesc[b] & 0xff) != BLK_NONE;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t o = (size_t)(3 * b + e) * ld + p;
            if (zo) zo[o] = split ? z[o] : x[o];
            if (uo) uo[o] = split ? u[o] : 0.0;
        }
    }
}

#endif  // ADMMB_ITERATE_ONLY

}  // namespace admmb
