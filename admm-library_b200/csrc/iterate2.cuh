// iterate2.cuh -- the decoupled fast iteration with TWO problems per thread.
//
// Why: (1) every broadcast LDS.128 of the shared factor now feeds two problems' FMAs, halving the load on the
// LSU data pipe, the busiest unit of the one-problem kernel at full width; (2) a lone warp per SM sub-partition
// spends ~30 % of its time in fixed-latency dependency waits -- the second problem is an independent instruction
// stream that fills them (the latency-bound tail of a solve is what bounds time-to-tolerance).
// Thread t of a CTA owns columns base+t and base+T+t (both coalesced).  Scope: shared factor staged in smem,
// decoupled model, "states unsplit / controls split" pattern, no affine term, no linear cost, shared parameter
// table (all five benchmark configurations).  Per-accumulator operation order is unchanged: bit-identical.
#pragma once
#include "kernels.cuh"

namespace admmb {

template <bool ADAPT>
__device__ __forceinline__ void admm_iteration_dec2(const IterParams &P, const size_t (&p)[2], const bool (&act)[2],
                                                    const FacRef<true> F, const int *bdesc, const uint32_t par_sbase,
                                                    const double (&rho)[2], const double (&sigma)[2], double (&nr)[2][5])
{
    constexpr bool FSH = true, FSMEM = true;
    const int N = P.N;
    const size_t ld = P.ld;
    const ptrdiff_t ld1 = (ptrdiff_t)ld, ld2 = 2 * (ptrdiff_t)ld, ld3 = 3 * (ptrdiff_t)ld;
    double rinv[2];
    double *zp[2], *up[2], *dp[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        rinv[j] = 1.0 / rho[j];
        zp[j] = P.z + p[j]; up[j] = P.u + p[j]; dp[j] = P.d + p[j];
    }
    auto load_par = [&](int b, double (&pr)[8]) {
        const uint32_t a = par_sbase + (uint32_t)b * 64u;
        const double2 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32), w = lds128(a + 48);
        pr[0] = x.x; pr[1] = x.y; pr[2] = y.x; pr[3] = y.y; pr[4] = z.x; pr[5] = z.y; pr[6] = w.x; pr[7] = w.y;
    };
    auto rt_terminal = [&](int j, int b, double (&t)[3]) {
        const int de = bdesc[b];
        if ((de & 0xff) != BLK_NONE) {
            const size_t r0 = (size_t)(de >> 8) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                double uu = ADMMB_LD(up[j] + (r0 + e) * ld);
                if (ADAPT) uu = uu * sigma[j];
                t[e] = ADMMB_LD(zp[j] + (r0 + e) * ld) - uu;
            }
        } else {
            t[0] = 0.0; t[1] = 0.0; t[2] = 0.0;
        }
    };

    // ---------------- backward sweep (g split in-plane / cross-track, ping-pong A/B, prefetch distance 2)
    double giA[2][4], gcA[2][2], giB[2][4], gcB[2][2];
    double zb[2][2][3], ub[2][2][3];                   // [slot][problem][e]
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        double t0[3], t1[3];
        rt_terminal(j, 3 * N, t0);
        rt_terminal(j, 3 * N + 1, t1);
        giA[j][0] = t0[0]; giA[j][1] = t0[1]; gcA[j][0] = t0[2]; giA[j][2] = t1[0]; giA[j][3] = t1[1]; gcA[j][1] = t1[2];
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (N - 1 - t >= 0) {
                const double *z0 = zp[j] + (ptrdiff_t)(N - 1 - t) * ld3, *u0 = up[j] + (ptrdiff_t)(N - 1 - t) * ld3;
                zb[t][j][0] = ADMMB_LD(z0); zb[t][j][1] = ADMMB_LD(z0 + ld1); zb[t][j][2] = ADMMB_LD(z0 + ld2);
                ub[t][j][0] = ADMMB_LD(u0); ub[t][j][1] = ADMMB_LD(u0 + ld1); ub[t][j][2] = ADMMB_LD(u0 + ld2);
            }
    }
    ptrdiff_t off_l = (ptrdiff_t)(N - 3) * ld3;        // rows of stage k-2 (loads), relative to zp / up
    ptrdiff_t off_s = (ptrdiff_t)(N - 1) * ld3;        // rows of stage k (d stores)
    auto bwd_stage = [&](const int k, double (&zc)[2][3], double (&uc)[2][3], const double (&gi)[2][4],
                         const double (&gc)[2][2], double (&pi)[2][4], double (&pc)[2][2]) {
        double ra[2][3], dj[2][3];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const double uu = ADAPT ? uc[j][e] * sigma[j] : uc[j][e];
                ra[j][e] = zc[j][e] - uu;
            }
            if (k >= 2) {
                const double *zl = zp[j] + off_l, *ul = up[j] + off_l;
                zc[j][0] = ADMMB_LD(zl); zc[j][1] = ADMMB_LD(zl + ld1); zc[j][2] = ADMMB_LD(zl + ld2);
                uc[j][0] = ADMMB_LD(ul); uc[j][1] = ADMMB_LD(ul + ld1); uc[j][2] = ADMMB_LD(ul + ld2);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) pi[j][i] = 0.0;
            pc[j][0] = 0.0; pc[j][1] = 0.0;
        }
        off_l -= ld3;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double h[2], er[4];
            dec_ld<FSH, FSMEM, 2>(F, k, D_HIN + 2 * r, h);
            dec_ld<FSH, FSMEM, 4>(F, k, D_EIN + 4 * r, er);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double acc = h[0] * ra[j][0];
                acc = fma(h[1], ra[j][1], acc);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc = fma(er[i], gi[j][i], acc);
                dj[j][r] = acc;
            }
        }
        {
            double h[2], er[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_HC, h);
            dec_ld<FSH, FSMEM, 2>(F, k, D_EC, er);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double acc = h[0] * ra[j][2];
                acc = fma(er[0], gc[j][0], acc);
                acc = fma(er[1], gc[j][1], acc);
                dj[j][2] = acc;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (act[j]) {
                double *ds = dp[j] + off_s;
                ADMMB_ST(ds, dj[j][0]); ADMMB_ST(ds + ld1, dj[j][1]); ADMMB_ST(ds + ld2, dj[j][2]);
            }
        off_s -= ld3;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double kr[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_KIN + 4 * r, kr);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) pi[j][i] = fma(kr[i], ra[j][r], pi[j][i]);
        }
        {
            double kr[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_KC, kr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pc[j][0] = fma(kr[0], ra[j][2], pc[j][0]);
                pc[j][1] = fma(kr[1], ra[j][2], pc[j][1]);
            }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            double ar[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_ACLIN + 4 * l, ar);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) pi[j][i] = fma(ar[i], gi[j][l], pi[j][i]);
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            double ar[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_ACLC + 2 * l, ar);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pc[j][0] = fma(ar[0], gc[j][l], pc[j][0]);
                pc[j][1] = fma(ar[1], gc[j][l], pc[j][1]);
            }
        }
    };
    {
        int k = N - 1;
        for (; k >= 1; k -= 2) {
            bwd_stage(k, zb[0], ub[0], giA, gcA, giB, gcB);
            bwd_stage(k - 1, zb[1], ub[1], giB, gcB, giA, gcA);
        }
        if (k == 0) bwd_stage(0, zb[0], ub[0], giA, gcA, giB, gcB);
    }

    // ---------------- forward sweep fused with prox / dual ascent / norms
    double rr[2] = {0.0, 0.0}, ss[2] = {0.0, 0.0}, xx[2] = {0.0, 0.0}, zz[2] = {0.0, 0.0}, uu[2] = {0.0, 0.0};
    double siA[2][4], scA[2][2], siB[2][4], scB[2][2], db[2][2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double *s0 = P.s0 + p[j];
        siA[j][0] = s0[0]; siA[j][1] = s0[ld]; scA[j][0] = s0[2 * ld];
        siA[j][2] = s0[3 * ld]; siA[j][3] = s0[4 * ld]; scA[j][1] = s0[5 * ld];
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (t < N) {
                const double *z0 = zp[j] + (ptrdiff_t)t * ld3, *u0 = up[j] + (ptrdiff_t)t * ld3, *d0 = dp[j] + (ptrdiff_t)t * ld3;
                zb[t][j][0] = ADMMB_LD(z0); zb[t][j][1] = ADMMB_LD(z0 + ld1); zb[t][j][2] = ADMMB_LD(z0 + ld2);
                ub[t][j][0] = ADMMB_LD(u0); ub[t][j][1] = ADMMB_LD(u0 + ld1); ub[t][j][2] = ADMMB_LD(u0 + ld2);
                db[t][j][0] = ADMMB_LD(d0); db[t][j][1] = ADMMB_LD(d0 + ld1); db[t][j][2] = ADMMB_LD(d0 + ld2);
            }
    }
    ptrdiff_t off_f = 2 * ld3;                          // rows of stage k+2 (loads)
    ptrdiff_t off_w = 0;                                // rows of stage k (z,u stores)
    auto fwd_stage = [&](const int k, double (&zc)[2][3], double (&uc)[2][3], double (&dc)[2][3], const double (&si)[2][4],
                         const double (&sc)[2][2], double (&ni)[2][4], double (&nc)[2][2]) {
        double a[2][3], zo[2][3], uo[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double kr[4];
            dec_ld<FSH, FSMEM, 4>(F, k, D_KIN + 4 * r, kr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double acc = dc[j][r];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc = fma(kr[i], si[j][i], acc);
                a[j][r] = acc;
            }
        }
        {
            double kr[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_KC, kr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double acc = dc[j][2];
                acc = fma(kr[0], sc[j][0], acc);
                acc = fma(kr[1], sc[j][1], acc);
                a[j][2] = acc;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 3; ++e) { zo[j][e] = zc[j][e]; uo[j][e] = ADAPT ? uc[j][e] * sigma[j] : uc[j][e]; }
            if (k + 2 < N) {
                const double *zf = zp[j] + off_f, *uf = up[j] + off_f, *df = dp[j] + off_f;
                zc[j][0] = ADMMB_LD(zf); zc[j][1] = ADMMB_LD(zf + ld1); zc[j][2] = ADMMB_LD(zf + ld2);
                uc[j][0] = ADMMB_LD(uf); uc[j][1] = ADMMB_LD(uf + ld1); uc[j][2] = ADMMB_LD(uf + ld2);
                dc[j][0] = ADMMB_LD(df); dc[j][1] = ADMMB_LD(df + ld1); dc[j][2] = ADMMB_LD(df + ld2);
            }
        }
        off_f += ld3;
        {
            const int b = 3 * k + 2;
            const int type = bdesc[b] & 0xff;
            double pr[8];
            load_par(b, pr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double v[3], zn[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const double xh = fma(P.alpha, a[j][e], P.oma * zo[j][e]);
                    v[e] = xh + uo[j][e];
                }
                prox_block_dev(type, [&](int q) { return pr[q]; }, rinv[j], v, zn);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const double un = v[e] - zn[e];
                    const double dr = a[j][e] - zn[e];
                    const double dz = zn[e] - zo[j][e];
                    rr[j] = fma(dr, dr, rr[j]);
                    ss[j] = fma(dz, dz, ss[j]);
                    xx[j] = fma(a[j][e], a[j][e], xx[j]);
                    zz[j] = fma(zn[e], zn[e], zz[j]);
                    uu[j] = fma(un, un, uu[j]);
                    if (act[j]) {
                        ADMMB_ST(zp[j] + off_w + e * ld1, zn[e]);
                        ADMMB_ST(up[j] + off_w + e * ld1, un);
                    }
                }
            }
        }
        off_w += ld3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double ar[4], br[2];
            dec_ld<FSH, FSMEM, 4>(F, k, D_AIN + 4 * i, ar);
            dec_ld<FSH, FSMEM, 2>(F, k, D_BIN + 2 * i, br);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double acc = ar[0] * si[j][0];
                acc = fma(ar[1], si[j][1], acc);
                acc = fma(ar[2], si[j][2], acc);
                acc = fma(ar[3], si[j][3], acc);
                acc = fma(br[0], a[j][0], acc);
                acc = fma(br[1], a[j][1], acc);
                ni[j][i] = acc;
            }
        }
        {
            double bc[2];
            dec_ld<FSH, FSMEM, 2>(F, k, D_BC, bc);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double ar[2];
                dec_ld<FSH, FSMEM, 2>(F, k, D_AC + 2 * i, ar);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double acc = ar[0] * sc[j][0];
                    acc = fma(ar[1], sc[j][1], acc);
                    acc = fma(bc[i], a[j][2], acc);
                    nc[j][i] = acc;
                }
            }
        }
    };
    {
        int k = 0;
        for (; k + 1 < N; k += 2) {
            fwd_stage(k, zb[0], ub[0], db[0], siA, scA, siB, scB);
            fwd_stage(k + 1, zb[1], ub[1], db[1], siB, scB, siA, scA);
        }
        if (k < N) fwd_stage(k, zb[0], ub[0], db[0], siA, scA, siB, scB);
    }
    const bool s_in_A = (N & 1) == 0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int b = 3 * N + t;
        const int de = bdesc[b];
        if ((de & 0xff) == BLK_NONE) continue;
        const size_t r0 = (size_t)(de >> 8) * 3;
        double pr[8];
        load_par(b, pr);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double xb[3] = {s_in_A ? siA[j][2 * t] : siB[j][2 * t], s_in_A ? siA[j][2 * t + 1] : siB[j][2 * t + 1],
                                  s_in_A ? scA[j][t] : scB[j][t]};
            double v[3], zn[3], zo[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                zo[e] = ADMMB_LD(zp[j] + (r0 + e) * ld);
                const double uv = ADMMB_LD(up[j] + (r0 + e) * ld);
                const double uo = ADAPT ? uv * sigma[j] : uv;
                const double xh = fma(P.alpha, xb[e], P.oma * zo[e]);
                v[e] = xh + uo;
            }
            prox_block_dev(de & 0xff, [&](int q) { return pr[q]; }, rinv[j], v, zn);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const double un = v[e] - zn[e];
                const double dr = xb[e] - zn[e];
                const double dz = zn[e] - zo[e];
                rr[j] = fma(dr, dr, rr[j]);
                ss[j] = fma(dz, dz, ss[j]);
                xx[j] = fma(xb[e], xb[e], xx[j]);
                zz[j] = fma(zn[e], zn[e], zz[j]);
                uu[j] = fma(un, un, uu[j]);
                if (act[j]) {
                    ADMMB_ST(zp[j] + (r0 + e) * ld, zn[e]);
                    ADMMB_ST(up[j] + (r0 + e) * ld, un);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) { nr[j][0] = rr[j]; nr[j][1] = ss[j]; nr[j][2] = xx[j]; nr[j][3] = zz[j]; nr[j][4] = uu[j]; }
}

// persistent launch, two problems per thread: columns blockIdx.x * 2T + threadIdx.x and + T
template <bool ADAPT>
__global__ void __launch_bounds__(256, 1) k_admm_iterate2(const __grid_constant__ IterParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *facS = reinterpret_cast<double *>(smem_raw + 16);
    double *parS = facS + (size_t)FD * P.N;
    int *bdS = reinterpret_cast<int *>(parS + 8 * P.nb);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t fac_sbase = (uint32_t)__cvta_generic_to_shared(facS);
    const uint32_t par_sbase = (uint32_t)__cvta_generic_to_shared(parS);
    const uint32_t bd_bytes = (uint32_t)(((P.nb + 3) / 4) * 16);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, bd_bytes + (uint32_t)(FD * P.N * 8) + (uint32_t)(8 * P.nb * 8));
        bulk_g2s(fac_sbase, P.fac_dec, (uint32_t)(FD * P.N * 8), mbar);
        bulk_g2s(par_sbase, P.par, (uint32_t)(8 * P.nb * 8), mbar);
        bulk_g2s((uint32_t)__cvta_generic_to_shared(bdS), P.bdesc, bd_bytes, mbar);
    }
    __syncthreads();
    mbar_wait(mbar, 0);

    const int T = blockDim.x;
    const int t0 = blockIdx.x * 2 * T + threadIdx.x;
    if (t0 >= P.n_active) return;
    size_t p[2] = {(size_t)t0, (size_t)(t0 + T < P.n_active ? t0 + T : t0)};
    int st[2], it[2];
    double rho[2], sigma[2], r_norm[2], s_norm[2], eps_pri[2], eps_dual[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        st[j] = P.status[p[j]];
        it[j] = P.iters[p[j]];
        rho[j] = P.rho[p[j]];
        sigma[j] = ADAPT ? P.usc[p[j]] : 1.0;
        r_norm[j] = s_norm[j] = eps_pri[j] = eps_dual[j] = 0.0;
    }
    const bool own1 = t0 + T < P.n_active;
    if (!own1) st[1] = ST_MAX_ITER;                    // no second problem: a finished dummy (aliases problem 0, never stored)
    bool was_running[2] = {st[0] == ST_RUNNING, st[1] == ST_RUNNING};
    if (!was_running[0] && !was_running[1]) return;

    FacRef<true> F;
    F.base = facS; F.ld = P.ld; F.sbase = fac_sbase;
    for (int cnt = 0; cnt < P.chunk && (st[0] == ST_RUNNING || st[1] == ST_RUNNING); ++cnt) {
        const bool act[2] = {st[0] == ST_RUNNING, st[1] == ST_RUNNING};
        double nr[2][5];
        admm_iteration_dec2<ADAPT>(P, p, act, F, bdS, par_sbase, rho, sigma, nr);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!act[j]) continue;
            ++it[j];
            sigma[j] = 1.0;
            r_norm[j] = sqrt(nr[j][0]);
            s_norm[j] = rho[j] * sqrt(nr[j][1]);
            const double nx = sqrt(nr[j][2]), nz = sqrt(nr[j][3]);
            eps_pri[j] = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
            eps_dual[j] = fma(P.reltol, rho[j] * sqrt(nr[j][4]), P.sqrtn_abs);
            if (P.hist) {
                const size_t h = (size_t)(it[j] - 1) * P.hist_ld + (P.orig ? (size_t)P.orig[p[j]] : p[j]);
                P.hist[h] = r_norm[j];
                P.hist[h + P.hist_stride] = s_norm[j];
                P.hist[h + 2 * P.hist_stride] = eps_pri[j];
                P.hist[h + 3 * P.hist_stride] = eps_dual[j];
                P.hist[h + 4 * P.hist_stride] = rho[j];
            }
            if (!(isfinite(r_norm[j]) && isfinite(s_norm[j]))) { st[j] = ST_NAN; continue; }
            if (r_norm[j] < eps_pri[j] && s_norm[j] < eps_dual[j]) { st[j] = ST_CONVERGED; continue; }
            if (ADAPT && (it[j] % P.every) == 0 && it[j] < P.max_iter && (P.until <= 0 || it[j] <= P.until)) {
                if (r_norm[j] > P.mu * s_norm[j]) {
                    if (!(rho[j] * P.tau > RHO_MAX)) { rho[j] = rho[j] * P.tau; sigma[j] = P.inv_tau; }
                } else if (s_norm[j] > P.mu * r_norm[j]) {
                    if (!(rho[j] * P.inv_tau < RHO_MIN)) { rho[j] = rho[j] * P.inv_tau; sigma[j] = P.tau; }
                }
            }
            if (it[j] >= P.max_iter) st[j] = ST_MAX_ITER;
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (!was_running[j]) continue;
        const size_t q = p[j];
        P.iters[q] = it[j];
        P.rho[q] = rho[j];
        if (ADAPT) P.usc[q] = sigma[j];
        P.status[q] = st[j];
        P.fin[q] = r_norm[j];
        P.fin[q + P.ld] = s_norm[j];
        P.fin[q + 2 * P.ld] = eps_pri[j];
        P.fin[q + 3 * P.ld] = eps_dual[j];
    }
}

}  // namespace admmb
