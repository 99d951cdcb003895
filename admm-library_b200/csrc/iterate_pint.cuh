// iterate_pint.cuh -- the PARALLEL-IN-TIME persistent kernel: the 2N-step dependent chain of the
// Riccati x-update is cut into C chunks of stages that C warps sweep AT THE SAME TIME.
//
// The two sweeps are linear recurrences over the stages,
//     backward   g_k = K_k' ra_k + Acl_k' g_{k+1},          d_k = Hinv_k ra_k + E_k g_{k+1}
//     forward    s_{k+1} = A_k s_k + B_k a_k,  a_k = d_k + K_k s_k     (homogeneous part: s_{k+1} = Acl_k s_k)
// so a chunk [k0, k1) can run its sweep from a ZERO incoming value (g_{k1} = 0 resp. s_{k0} = 0) -- the particular
// solution -- and be corrected afterwards by the homogeneous response to the true incoming value:
//     d_k = d~_k + W_k g_{k1},   W_k = E_k  Acl_{k+1}' ... Acl_{k1-1}'          a_k = a~_k + V_k s_{k0},   V_k = K_k Acl_{k-1} ... Acl_{k0}
// W_k, V_k and the chunk-to-chunk transitions T_c, S_c depend on the shared factor only and are computed once per
// factor (k_pint_pack).  The true incoming values follow from the chunks' outgoing ones by C - 1 small matrix-vector
// products.  Sequential depth of an iteration: 2 N / C stages + 2 C boundary products instead of 2 N stages.
//
// A CTA of eight warps owns a resident tile of TW <= 32 problems (lane = problem; z, u, d, a live in shared memory for
// the whole launch, as in iterate_wg.cuh).  Warp c does EVERYTHING for its chunk of stages -- backward sweep, forward
// sweep, relaxation / prox / dual ascent and the partial norm sums of its control blocks -- so the only hand-overs are
// four CTA barriers per iteration (chunk boundaries twice, norm partials, roots).
//
// Precision class: FP64 throughout, but NOT the oracle's operation order (the superposition re-associates the
// recurrences, the five norms are summed per chunk first).  Results agree with oracle/admm_ocp_cpu.c to ~1e-12 relative
// (tests hold them to the north_star's 1e-9 and the iteration counts to +-1 on a few threshold cases); the kernel is
// therefore OPT-IN (opts.kernel = ADMMB_KERNEL_PINT) and never chosen automatically: the default path stays bit-exact.
//
// Scope (as iterate_wg.cuh): shared decoupled factor, "states unsplit, controls split" pattern, no affine term, no
// linear cost, shared parameter table.
#pragma once
#ifdef WG_TIMING
#include <cstdio>
#endif
#include "iterate_wg.cuh"

namespace admmb {

#ifndef PT_WARPS_N
#define PT_WARPS_N 8
#endif
constexpr int PT_WARPS = PT_WARPS_N;
constexpr int PT_STAGE = 20;      // per stage: Wi[2][4] Wc[2] Vi[2][4] Vc[2]
constexpr int PT_CHUNK = 40;      // per chunk: Tin[4][4] Tc[2][2] Sin[4][4] Sc[2][2]

__host__ __device__ inline int pint_chunks(int N) { return N < PT_WARPS ? N : PT_WARPS; }
__host__ __device__ inline int pint_k0(int c, int N, int C) { return (int)(((long long)c * N) / C); }
__host__ __device__ inline size_t pint_table_doubles(int N) { return (size_t)PT_STAGE * N + (size_t)PT_CHUNK * PT_WARPS; }

// once per factor: the correction matrices of every stage and the transitions of every chunk (one thread; N <= ~1000)
__global__ void k_pint_pack(int N, const double *__restrict__ fd, double *__restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int C = pint_chunks(N);
    double *pst = out, *pch = out + (size_t)PT_STAGE * N;
    for (int c = 0; c < C; ++c) {
        const int k0 = pint_k0(c, N, C), k1 = pint_k0(c + 1, N, C);
        // backward: T = Acl_k' ... Acl_{k1-1}' (T(k1) = I), W_k = E_k T(k+1)
        double T[4][4], Tc[2][2];
        for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) T[i][m] = i == m ? 1.0 : 0.0;
        for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) Tc[i][m] = i == m ? 1.0 : 0.0;
        for (int k = k1 - 1; k >= k0; --k) {
            const double *f = fd + (size_t)k * FD;
            double *o = pst + (size_t)k * PT_STAGE;
            for (int j = 0; j < 2; ++j)
                for (int m = 0; m < 4; ++m) {
                    double acc = 0.0;
                    for (int i = 0; i < 4; ++i) acc = fma(f[D_EIN + 4 * j + i], T[i][m], acc);
                    o[4 * j + m] = acc;
                }
            for (int m = 0; m < 2; ++m) o[8 + m] = fma(f[D_EC + 1], Tc[1][m], f[D_EC] * Tc[0][m]);
            double Tn[4][4], Tcn[2][2];
            for (int i = 0; i < 4; ++i)
                for (int m = 0; m < 4; ++m) {
                    double acc = 0.0;
                    for (int l = 0; l < 4; ++l) acc = fma(f[D_ACLIN + 4 * l + i], T[l][m], acc);     // (Acl' T)[i][m]
                    Tn[i][m] = acc;
                }
            for (int i = 0; i < 2; ++i)
                for (int m = 0; m < 2; ++m) Tcn[i][m] = fma(f[D_ACLC + 2 + i], Tc[1][m], f[D_ACLC + i] * Tc[0][m]);
            for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) T[i][m] = Tn[i][m];
            for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) Tc[i][m] = Tcn[i][m];
        }
        double *oc = pch + (size_t)c * PT_CHUNK;
        for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) oc[4 * i + m] = T[i][m];
        for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) oc[16 + 2 * i + m] = Tc[i][m];
        // forward: S = Acl_{k-1} ... Acl_{k0} (S(k0) = I), V_k = K_k S(k)
        double S[4][4], Sc[2][2];
        for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) S[i][m] = i == m ? 1.0 : 0.0;
        for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) Sc[i][m] = i == m ? 1.0 : 0.0;
        for (int k = k0; k < k1; ++k) {
            const double *f = fd + (size_t)k * FD;
            double *o = pst + (size_t)k * PT_STAGE + 10;
            for (int j = 0; j < 2; ++j)
                for (int m = 0; m < 4; ++m) {
                    double acc = 0.0;
                    for (int i = 0; i < 4; ++i) acc = fma(f[D_KIN + 4 * j + i], S[i][m], acc);
                    o[4 * j + m] = acc;
                }
            for (int m = 0; m < 2; ++m) o[8 + m] = fma(f[D_KC + 1], Sc[1][m], f[D_KC] * Sc[0][m]);
            double Sn[4][4], Scn[2][2];
            for (int l = 0; l < 4; ++l)
                for (int m = 0; m < 4; ++m) {
                    double acc = 0.0;
                    for (int i = 0; i < 4; ++i) acc = fma(f[D_ACLIN + 4 * l + i], S[i][m], acc);     // (Acl S)[l][m]
                    Sn[l][m] = acc;
                }
            for (int l = 0; l < 2; ++l)
                for (int m = 0; m < 2; ++m) Scn[l][m] = fma(f[D_ACLC + 2 * l + 1], Sc[1][m], f[D_ACLC + 2 * l] * Sc[0][m]);
            for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) S[i][m] = Sn[i][m];
            for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) Sc[i][m] = Scn[i][m];
        }
        for (int i = 0; i < 4; ++i) for (int m = 0; m < 4; ++m) oc[20 + 4 * i + m] = S[i][m];
        for (int i = 0; i < 2; ++i) for (int m = 0; m < 2; ++m) oc[36 + 2 * i + m] = Sc[i][m];
    }
}

// Shared-memory accesses of this kernel go through ordinary pointers: every tile array (z, u, d, a) is read and written by
// ONE warp between two CTA barriers, the factor and the tables are read-only, so the compiler is free to hoist the loads
// of a stage above the arithmetic of the previous one (iterate_wg.cuh pins the order with volatile asm instead, because
// there other warps write what a warp reads).  `a` is a 32-bit shared-window address, `b` the window address of pt_smem[0].
#define PT_PTR(a) (pt_smem + ((a) - sm0))
#define pt_ld(a) (*reinterpret_cast<const double *>(PT_PTR(a)))
#define pt_st(a, v) (*reinterpret_cast<double *>(PT_PTR(a)) = (v))
#define pt_ld2(a, v0_, v1_) do { const double2 t2_ = *reinterpret_cast<const double2 *>(PT_PTR(a)); (v0_) = t2_.x; (v1_) = t2_.y; } while (0)
#define pt_ld4(a, r) do { pt_ld2((a), (r)[0], (r)[1]); pt_ld2((a) + 16, (r)[2], (r)[3]); } while (0)

struct PintLayout {               // byte offsets inside the dynamic shared memory of one CTA
    int TW, rz, rd, rb, rn;
    size_t fac, pst, pch, par, typ, z, u, d, a, s0, bnd, nrm, total;
};
__host__ __device__ inline PintLayout pint_layout(int N, int rows_zu, int TW)
{
    PintLayout L;
    L.TW = TW;
    L.rz = rows_zu | 1;
    L.rd = (3 * N) | 1;
    L.rb = (12 * PT_WARPS) | 1;                  // per chunk: outgoing g~ (6) and s~ (6)
    L.rn = (5 * PT_WARPS + 5) | 1;               // per chunk: five partial sums; then the five roots
    const int nsb = rows_zu / 3;
    size_t o = 16;                               // [0, 8): mbarrier of the bulk copies
    L.fac = o; o += sizeof(double) * FD * (size_t)N;
    L.pst = o; o += sizeof(double) * PT_STAGE * (size_t)N;
    L.pch = o; o += sizeof(double) * PT_CHUNK * PT_WARPS;
    L.par = o; o += sizeof(double) * 8 * (size_t)nsb;
    L.typ = o; o += sizeof(int) * (size_t)((nsb + 3) / 4 * 4);
    const size_t col = sizeof(double) * (size_t)TW;
    L.z = o; o += col * L.rz;
    L.u = o; o += col * L.rz;
    L.d = o; o += col * L.rd;
    L.a = o; o += col * L.rd;
    L.s0 = o; o += col * 7;
    L.bnd = o; o += col * L.rb;
    L.nrm = o; o += col * L.rn;
    L.total = (o + 15) / 16 * 16;
    return L;
}

#ifdef WG_TIMING
// developer build: SM-clock stamps of one iteration (the 8th of the launch) of CTA 0, printed when the launch ends
__device__ long long pt_tlog[PT_WARPS][10];
#define PT_T(idx) do { if (blockIdx.x == 0 && (tid & 31) == 0 && cnt == 7) pt_tlog[warp][idx] = clock64(); } while (0)
#else
#define PT_T(idx) do { } while (0)
#endif

// TI: A_k, B_k do not depend on the stage (checked at upload): kept in registers for the whole launch
template <bool ADAPT, bool TI>
__global__ void __launch_bounds__(PT_WARPS * 32, 1)
k_admm_iterate_pint(const __grid_constant__ IterParams P, const double *__restrict__ ptab, const int TW)
{
    extern __shared__ __align__(16) unsigned char pt_smem[];
    const int N = P.N;
    const int C = pint_chunks(N);
    const PintLayout L = pint_layout(N, P.rows_zu, TW);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int lane = min(tid & 31, TW - 1);              // lanes beyond the tile width are clones of its last lane
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(pt_smem);
    const uint32_t mbar = sm0;
    const uint32_t fac_s = sm0 + (uint32_t)L.fac, pst_s = sm0 + (uint32_t)L.pst, pch_s = sm0 + (uint32_t)L.pch;
    const uint32_t par_s = sm0 + (uint32_t)L.par, typ_s = sm0 + (uint32_t)L.typ;
    const uint32_t zs = sm0 + (uint32_t)L.z + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t us = sm0 + (uint32_t)L.u + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t ds = sm0 + (uint32_t)L.d + (uint32_t)(lane * L.rd) * 8u;
    const uint32_t as = sm0 + (uint32_t)L.a + (uint32_t)(lane * L.rd) * 8u;
    const uint32_t s0s = sm0 + (uint32_t)L.s0 + (uint32_t)(lane * 7) * 8u;
    const uint32_t bnd = sm0 + (uint32_t)L.bnd + (uint32_t)(lane * L.rb) * 8u;
    const uint32_t nrm = sm0 + (uint32_t)L.nrm + (uint32_t)(lane * L.rn) * 8u;
    const uint32_t roots = nrm + (uint32_t)(5 * PT_WARPS) * 8u;

    if (tid == 0) {
        mbar_init(mbar, 1);
        const uint32_t b0 = (uint32_t)(FD * N * 8), b1 = (uint32_t)(PT_STAGE * N * 8), b2 = (uint32_t)(PT_CHUNK * PT_WARPS * 8);
        mbar_expect_tx(mbar, b0 + b1 + b2);
        bulk_g2s(fac_s, P.fac_dec, b0, mbar);
        bulk_g2s(pst_s, ptab, b1, mbar);
        bulk_g2s(pch_s, ptab + (size_t)PT_STAGE * N, b2, mbar);
    }
    {
        double *parS = reinterpret_cast<double *>(pt_smem + L.par);
        int *typS = reinterpret_cast<int *>(pt_smem + L.typ);
        for (int b = tid; b < P.nb; b += PT_WARPS * 32) {
            const int de = P.bdesc[b];
            if ((de & 0xff) != BLK_NONE) {
                const int slot = de >> 8;
                typS[slot] = de & 0xff;
#pragma unroll
                for (int q = 0; q < 8; ++q) parS[8 * slot + q] = P.par[8 * b + q];
            }
        }
    }
    const int de_t0 = P.bdesc[3 * N], de_t1 = P.bdesc[3 * N + 1];      // terminal blocks (any type, maybe unsplit)
    __syncthreads();
    mbar_wait(mbar, 0);

    double Ai[4][4], Bi[4][2], Ac[2][2], Bc[2];
    if (TI) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            pt_ld4(fac_s + (D_AIN + 4 * r) * 8, Ai[r]);
            pt_ld2(fac_s + (D_BIN + 2 * r) * 8, Bi[r][0], Bi[r][1]);
        }
        pt_ld2(fac_s + D_AC * 8, Ac[0][0], Ac[0][1]);
        pt_ld2(fac_s + (D_AC + 2) * 8, Ac[1][0], Ac[1][1]);
        pt_ld2(fac_s + D_BC * 8, Bc[0], Bc[1]);
    }
    const bool chunk_w = warp < C;
    const int k0 = chunk_w ? pint_k0(warp, N, C) : 0, k1 = chunk_w ? pint_k0(warp + 1, N, C) : 0;
    const bool first = warp == 0, last = warp == C - 1;

    const size_t ld = P.ld;
    const int ntiles = (P.n_active + TW - 1) / TW;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TW + lane;
        const bool valid = t < P.n_active;
        const size_t p = (size_t)t;
        const bool live = valid && P.status[t] == ST_RUNNING;
        if (!__syncthreads_or(live)) continue;

        // ---- tile in
        if (valid) {
#pragma unroll 4
            for (int r = warp; r < P.rows_zu; r += PT_WARPS) {
                pt_st(zs + 8u * r, __ldcg(P.z + (size_t)r * ld + p));
                pt_st(us + 8u * r, __ldcg(P.u + (size_t)r * ld + p));
            }
#pragma unroll 4
            for (int r = warp; r < 3 * N; r += PT_WARPS) pt_st(ds + 8u * r, __ldcg(P.d + (size_t)r * ld + p));
            if (warp == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) pt_st(s0s + 8u * i, P.s0[(size_t)i * ld + p]);
            }
        }
        double rho = live ? P.rho[p] : 1.0;
        double sigma = (ADAPT && live) ? P.usc[p] : 1.0;
        int it = live ? P.iters[p] : 0;
        int st = live ? ST_RUNNING : ST_MAX_ITER;
        double r_norm = 0.0, s_norm = 0.0, eps_pri = 0.0, eps_dual = 0.0;
        __syncthreads();

        for (int cnt = 0; cnt < P.chunk; ++cnt) {
            const bool run = st == ST_RUNNING;
            if (!__any_sync(0xffffffffu, run)) break;    // lane = problem in every warp: all warps hold the same flags
            const double rinv = 1.0 / rho;
            double si[4] = {0.0, 0.0, 0.0, 0.0}, sc0 = 0.0, sc1 = 0.0;      // outgoing state of this chunk's forward sweep
            PT_T(0);

            if (chunk_w) {
                // ================= backward sweep of this chunk, from g = 0 (the last chunk: from the terminal blocks)
                double gi[4] = {0.0, 0.0, 0.0, 0.0}, gc0 = 0.0, gc1 = 0.0;
                if (last) {
#pragma unroll
                    for (int tb = 0; tb < 2; ++tb) {
                        const int de = tb == 0 ? de_t0 : de_t1;
                        if ((de & 0xff) != BLK_NONE) {
                            const uint32_t o = (uint32_t)((de >> 8) * 3) * 8u;
                            double v[3];
#pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                double uu = pt_ld(us + o + 8u * e);
                                if (ADAPT) uu = uu * sigma;
                                v[e] = pt_ld(zs + o + 8u * e) - uu;
                            }
                            gi[2 * tb] = v[0]; gi[2 * tb + 1] = v[1];
                            if (tb == 0) gc0 = v[2]; else gc1 = v[2];
                        }
                    }
                }
#pragma unroll 1
                for (int k = k1 - 1; k >= k0; --k) {
                    const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                    double kr0[4], kr1[4], ac[4][4], h0[2], h1[2], e0[4], e1[4], kc[2], a0[2], a1[2], hc[2], ec[2], z[3], u[3];
                    const uint32_t o = (uint32_t)(3 * k) * 8u;
#pragma unroll
                    for (int e = 0; e < 3; ++e) { z[e] = pt_ld(zs + o + 8u * e); u[e] = pt_ld(us + o + 8u * e); }
                    pt_ld4(fk + D_KIN * 8, kr0);
                    pt_ld4(fk + (D_KIN + 4) * 8, kr1);
                    pt_ld2(fk + D_HIN * 8, h0[0], h0[1]);
                    pt_ld2(fk + (D_HIN + 2) * 8, h1[0], h1[1]);
                    pt_ld2(fk + D_KC * 8, kc[0], kc[1]);
                    pt_ld2(fk + D_HC * 8, hc[0], hc[1]);
#pragma unroll
                    for (int l = 0; l < 4; ++l) pt_ld4(fk + (D_ACLIN + 4 * l) * 8, ac[l]);
                    pt_ld4(fk + D_EIN * 8, e0);
                    pt_ld4(fk + (D_EIN + 4) * 8, e1);
                    pt_ld2(fk + D_ACLC * 8, a0[0], a0[1]);
                    pt_ld2(fk + (D_ACLC + 2) * 8, a1[0], a1[1]);
                    pt_ld2(fk + D_EC * 8, ec[0], ec[1]);
                    double ra[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double uu = ADAPT ? u[e] * sigma : u[e];
                        ra[e] = z[e] - uu;
                    }
                    double pi[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        pi[i] = kr0[i] * ra[0];
                        pi[i] = fma(kr1[i], ra[1], pi[i]);
                    }
                    double d0 = h0[0] * ra[0];
                    d0 = fma(h0[1], ra[1], d0);
                    double d1 = h1[0] * ra[0];
                    d1 = fma(h1[1], ra[1], d1);
                    double p0 = kc[0] * ra[2], p1 = kc[1] * ra[2], d2 = hc[0] * ra[2];
#pragma unroll
                    for (int l = 0; l < 4; ++l)
#pragma unroll
                        for (int i = 0; i < 4; ++i) pi[i] = fma(ac[l][i], gi[l], pi[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { d0 = fma(e0[i], gi[i], d0); d1 = fma(e1[i], gi[i], d1); }
                    p0 = fma(a0[0], gc0, p0); p1 = fma(a0[1], gc0, p1);
                    p0 = fma(a1[0], gc1, p0); p1 = fma(a1[1], gc1, p1);
                    d2 = fma(ec[0], gc0, d2);
                    d2 = fma(ec[1], gc1, d2);
                    if (run) { pt_st(ds + o, d0); pt_st(ds + o + 8, d1); pt_st(ds + o + 16, d2); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) gi[i] = pi[i];
                    gc0 = p0; gc1 = p1;
                }
                {   // outgoing g~ of this chunk
                    const uint32_t b = bnd + (uint32_t)(12 * warp) * 8u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) pt_st(b + 8u * i, gi[i]);
                    pt_st(b + 32, gc0); pt_st(b + 40, gc1);
                }
            }
            PT_T(1);
            __syncthreads();                                           // B1: every chunk's outgoing g~ is visible
            PT_T(2);

            if (chunk_w) {
                // ================= true g at this chunk's upper end: g(C-1) is true; g(j) = g~(j) + T_j g(j+1)
                double gb[4] = {0.0, 0.0, 0.0, 0.0}, gbc[2] = {0.0, 0.0};
                if (!last) {
                    {
                        const uint32_t b = bnd + (uint32_t)(12 * (C - 1)) * 8u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) gb[i] = pt_ld(b + 8u * i);
                        gbc[0] = pt_ld(b + 32); gbc[1] = pt_ld(b + 40);
                    }
#pragma unroll 1
                    for (int j = C - 2; j > warp; --j) {
                        const uint32_t b = bnd + (uint32_t)(12 * j) * 8u, tj = pch_s + (uint32_t)(j * PT_CHUNK) * 8u;
                        double n[4], nc[2], T[4][4], Tc[2][2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { pt_ld4(tj + (uint32_t)(4 * i) * 8u, T[i]); n[i] = pt_ld(b + 8u * i); }
#pragma unroll
                        for (int i = 0; i < 2; ++i) { pt_ld2(tj + (uint32_t)(16 + 2 * i) * 8u, Tc[i][0], Tc[i][1]); nc[i] = pt_ld(b + 32 + 8u * i); }
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int i = 0; i < 4; ++i) n[i] = fma(T[i][m], gb[m], n[i]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) nc[i] = fma(Tc[i][1], gbc[1], fma(Tc[i][0], gbc[0], nc[i]));
#pragma unroll
                        for (int i = 0; i < 4; ++i) gb[i] = n[i];
                        gbc[0] = nc[0]; gbc[1] = nc[1];
                    }
                }
                PT_T(3);
                // ================= forward sweep of this chunk from s = 0 (the first chunk: from s0); d_k gets its
                // correction W_k g on the way; a~_k is kept for the second pass
                if (first) {
                    si[0] = pt_ld(s0s); si[1] = pt_ld(s0s + 8); si[2] = pt_ld(s0s + 24); si[3] = pt_ld(s0s + 32);
                    sc0 = pt_ld(s0s + 16); sc1 = pt_ld(s0s + 40);
                }
#pragma unroll 1
                for (int k = k0; k < k1; ++k) {
                    const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u, o = (uint32_t)(3 * k) * 8u;
                    const uint32_t wk = pst_s + (uint32_t)(k * PT_STAGE) * 8u;
                    double dk[3], kn0[4], kn1[4], kc[2], an[4][4], bn[4][2], ac0[2], ac1[2], bc[2];
#pragma unroll
                    for (int e = 0; e < 3; ++e) dk[e] = pt_ld(ds + o + 8u * e);
                    pt_ld4(fk + D_KIN * 8, kn0);
                    pt_ld4(fk + (D_KIN + 4) * 8, kn1);
                    pt_ld2(fk + D_KC * 8, kc[0], kc[1]);
                    if (!last) {
                        double w0[4], w1[4], wc[2];
                        pt_ld4(wk, w0);
                        pt_ld4(wk + 32, w1);
                        pt_ld2(wk + 64, wc[0], wc[1]);
#pragma unroll
                        for (int m = 0; m < 4; ++m) { dk[0] = fma(w0[m], gb[m], dk[0]); dk[1] = fma(w1[m], gb[m], dk[1]); }
                        dk[2] = fma(wc[1], gbc[1], fma(wc[0], gbc[0], dk[2]));
                        if (run) { pt_st(ds + o, dk[0]); pt_st(ds + o + 8, dk[1]); pt_st(ds + o + 16, dk[2]); }
                    }
                    if (TI) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) an[r][i] = Ai[r][i];
                            bn[r][0] = Bi[r][0]; bn[r][1] = Bi[r][1];
                        }
                        ac0[0] = Ac[0][0]; ac0[1] = Ac[0][1]; ac1[0] = Ac[1][0]; ac1[1] = Ac[1][1]; bc[0] = Bc[0]; bc[1] = Bc[1];
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            pt_ld4(fk + (D_AIN + 4 * r) * 8, an[r]);
                            pt_ld2(fk + (D_BIN + 2 * r) * 8, bn[r][0], bn[r][1]);
                        }
                        pt_ld2(fk + D_AC * 8, ac0[0], ac0[1]);
                        pt_ld2(fk + (D_AC + 2) * 8, ac1[0], ac1[1]);
                        pt_ld2(fk + D_BC * 8, bc[0], bc[1]);
                    }
                    double a0 = dk[0], a1 = dk[1];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { a0 = fma(kn0[i], si[i], a0); a1 = fma(kn1[i], si[i], a1); }
                    double a2 = fma(kc[0], sc0, dk[2]);
                    a2 = fma(kc[1], sc1, a2);
                    pt_st(as + o, a0); pt_st(as + o + 8, a1); pt_st(as + o + 16, a2);
                    double ni[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        double acc = an[r][0] * si[0];
                        acc = fma(an[r][1], si[1], acc);
                        acc = fma(an[r][2], si[2], acc);
                        acc = fma(an[r][3], si[3], acc);
                        acc = fma(bn[r][0], a0, acc);
                        acc = fma(bn[r][1], a1, acc);
                        ni[r] = acc;
                    }
                    double n0 = ac0[0] * sc0;
                    n0 = fma(ac0[1], sc1, n0);
                    n0 = fma(bc[0], a2, n0);
                    double n1 = ac1[0] * sc0;
                    n1 = fma(ac1[1], sc1, n1);
                    n1 = fma(bc[1], a2, n1);
#pragma unroll
                    for (int r = 0; r < 4; ++r) si[r] = ni[r];
                    sc0 = n0; sc1 = n1;
                }
                {   // outgoing s~ of this chunk
                    const uint32_t b = bnd + (uint32_t)(12 * warp + 6) * 8u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) pt_st(b + 8u * i, si[i]);
                    pt_st(b + 32, sc0); pt_st(b + 40, sc1);
                }
            }
            PT_T(4);
            __syncthreads();                                           // B2: every chunk's outgoing s~ is visible
            PT_T(5);

            double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
            if (chunk_w) {
                // ================= true s at this chunk's lower end: s(0) = s0; s(j+1) = s~(j) + S_j s(j)
                double sb[4] = {0.0, 0.0, 0.0, 0.0}, sbc[2] = {0.0, 0.0};
                if (!first) {
                    {
                        const uint32_t b = bnd + 6u * 8u;              // chunk 0 started from the true s0
#pragma unroll
                        for (int i = 0; i < 4; ++i) sb[i] = pt_ld(b + 8u * i);
                        sbc[0] = pt_ld(b + 32); sbc[1] = pt_ld(b + 40);
                    }
#pragma unroll 1
                    for (int j = 1; j < warp; ++j) {
                        const uint32_t b = bnd + (uint32_t)(12 * j + 6) * 8u, sj = pch_s + (uint32_t)(j * PT_CHUNK + 20) * 8u;
                        double n[4], nc[2], S[4][4], Sc[2][2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { pt_ld4(sj + (uint32_t)(4 * i) * 8u, S[i]); n[i] = pt_ld(b + 8u * i); }
#pragma unroll
                        for (int i = 0; i < 2; ++i) { pt_ld2(sj + (uint32_t)(16 + 2 * i) * 8u, Sc[i][0], Sc[i][1]); nc[i] = pt_ld(b + 32 + 8u * i); }
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int i = 0; i < 4; ++i) n[i] = fma(S[i][m], sb[m], n[i]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) nc[i] = fma(Sc[i][1], sbc[1], fma(Sc[i][0], sbc[0], nc[i]));
#pragma unroll
                        for (int i = 0; i < 4; ++i) sb[i] = n[i];
                        sbc[0] = nc[0]; sbc[1] = nc[1];
                    }
                }
                PT_T(6);
                // ================= second pass over the chunk: a_k = a~_k + V_k s, then relaxation, prox, dual ascent and the
                // chunk's share of the five norms
                auto block = [&](int slot, const double (&x)[3]) {
                    const uint32_t pa = par_s + (uint32_t)slot * 64u, o = (uint32_t)(3 * slot) * 8u;
                    double pr[8], zo[3], uo[3];
                    int type;
                    pt_ld2(pa, pr[0], pr[1]); pt_ld2(pa + 16, pr[2], pr[3]);
                    pt_ld2(pa + 32, pr[4], pr[5]); pt_ld2(pa + 48, pr[6], pr[7]);
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(type) : "r"(typ_s + 4u * slot) : "memory");
#pragma unroll
                    for (int e = 0; e < 3; ++e) { zo[e] = pt_ld(zs + o + 8u * e); uo[e] = pt_ld(us + o + 8u * e); }
                    double v[3], zn[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double us_ = ADAPT ? uo[e] * sigma : uo[e];
                        const double xh = fma(P.alpha, x[e], P.oma * zo[e]);
                        v[e] = xh + us_;
                    }
                    prox_block_dev(type, [&](int q) { return pr[q]; }, rinv, v, zn);
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double un = v[e] - zn[e], dz = zn[e] - zo[e], dr = x[e] - zn[e];
                        if (run) { pt_st(zs + o + 8u * e, zn[e]); pt_st(us + o + 8u * e, un); }
                        rr = fma(dr, dr, rr);
                        ss = fma(dz, dz, ss);
                        xx = fma(x[e], x[e], xx);
                        zz = fma(zn[e], zn[e], zz);
                        uu = fma(un, un, uu);
                    }
                };
#pragma unroll 1
                for (int k = k0; k < k1; ++k) {
                    const uint32_t o = (uint32_t)(3 * k) * 8u, vk = pst_s + (uint32_t)(k * PT_STAGE + 10) * 8u;
                    double x[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) x[e] = pt_ld(as + o + 8u * e);
                    if (!first) {
                        double v0[4], v1[4], vc[2];
                        pt_ld4(vk, v0);
                        pt_ld4(vk + 32, v1);
                        pt_ld2(vk + 64, vc[0], vc[1]);
#pragma unroll
                        for (int m = 0; m < 4; ++m) { x[0] = fma(v0[m], sb[m], x[0]); x[1] = fma(v1[m], sb[m], x[1]); }
                        x[2] = fma(vc[1], sbc[1], fma(vc[0], sbc[0], x[2]));
                    }
                    block(k, x);                                       // fast pattern: control block k has compact index k
                }
                if (last) {
                    // terminal state s_N = s~ + S_{C-1} s, then the terminal blocks
                    double sN[6];
                    if (!first) {
                        const uint32_t sj = pch_s + (uint32_t)(warp * PT_CHUNK + 20) * 8u;
                        double S[4], n[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            pt_ld4(sj + (uint32_t)(4 * i) * 8u, S);
                            double acc = si[i];
#pragma unroll
                            for (int m = 0; m < 4; ++m) acc = fma(S[m], sb[m], acc);
                            n[i] = acc;
                        }
                        double t0, t1;
                        pt_ld2(sj + 16u * 8u, t0, t1);
                        const double c0 = fma(t1, sbc[1], fma(t0, sbc[0], sc0));
                        pt_ld2(sj + 18u * 8u, t0, t1);
                        const double c1 = fma(t1, sbc[1], fma(t0, sbc[0], sc1));
                        sN[0] = n[0]; sN[1] = n[1]; sN[2] = c0; sN[3] = n[2]; sN[4] = n[3]; sN[5] = c1;
                    } else {
                        sN[0] = si[0]; sN[1] = si[1]; sN[2] = sc0; sN[3] = si[2]; sN[4] = si[3]; sN[5] = sc1;
                    }
#pragma unroll
                    for (int tb = 0; tb < 2; ++tb) {
                        const int de = tb == 0 ? de_t0 : de_t1;
                        if ((de & 0xff) == BLK_NONE) continue;
                        const double x[3] = {sN[3 * tb], sN[3 * tb + 1], sN[3 * tb + 2]};
                        block(de >> 8, x);
                    }
                }
                const uint32_t b = nrm + (uint32_t)(5 * warp) * 8u;
                pt_st(b, rr); pt_st(b + 8, ss); pt_st(b + 16, xx); pt_st(b + 24, zz); pt_st(b + 32, uu);
            }
            PT_T(7);
            __syncthreads();                                           // B3: the chunks' partial sums are visible
            if (warp < 5) {                                            // one norm per warp: chunks in ascending order, root
                double acc = 0.0;
                for (int c = 0; c < C; ++c) acc += pt_ld(nrm + (uint32_t)(5 * c + warp) * 8u);
                pt_st(roots + 8u * warp, sqrt(acc));
            }
            __syncthreads();                                           // B4
            PT_T(8);
            if (run) {
                ++it;
                sigma = 1.0;
                r_norm = pt_ld(roots);
                s_norm = rho * pt_ld(roots + 8);
                const double nx = pt_ld(roots + 16), nz = pt_ld(roots + 24);
                eps_pri = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
                eps_dual = fma(P.reltol, rho * pt_ld(roots + 32), P.sqrtn_abs);
                if (P.hist && warp == 4) {
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
                        if (r_norm > P.mu * s_norm) {
                            if (!(rho * P.tau > RHO_MAX)) { rho = rho * P.tau; sigma = P.inv_tau; }
                        } else if (s_norm > P.mu * r_norm) {
                            if (!(rho * P.inv_tau < RHO_MIN)) { rho = rho * P.inv_tau; sigma = P.tau; }
                        }
                    }
                    if (it >= P.max_iter) st = ST_MAX_ITER;
                }
            }
            PT_T(9);
        }

#ifdef WG_TIMING
        if (blockIdx.x == 0 && tid == 0) {
            const char *names[10] = {"start", "bwd_done", "B1", "gb_done", "fwd_done", "B2", "sb_done", "pass2_done", "B4", "stop_done"};
            for (int w = 0; w < PT_WARPS; ++w)
                for (int i = 0; i < 10; ++i)
                    if (pt_tlog[w][i]) printf("ptt warp %d %-10s %lld\n", w, names[i], pt_tlog[w][i] - pt_tlog[0][0]);
        }
#endif
        // ---- tile out
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int r = warp; r < P.rows_zu; r += PT_WARPS) {
                __stcg(P.z + (size_t)r * ld + p, pt_ld(zs + 8u * r));
                __stcg(P.u + (size_t)r * ld + p, pt_ld(us + 8u * r));
            }
#pragma unroll 4
            for (int r = warp; r < 3 * N; r += PT_WARPS) __stcg(P.d + (size_t)r * ld + p, pt_ld(ds + 8u * r));
        }
        if (live && warp == 0) {
            P.iters[p] = it;
            P.rho[p] = rho;
            if (ADAPT) P.usc[p] = sigma;
            P.status[p] = st;
            P.fin[p] = r_norm;
            P.fin[p + ld] = s_norm;
            P.fin[p + 2 * ld] = eps_pri;
            P.fin[p + 3 * ld] = eps_dual;
        }
    }
}

}  // namespace admmb
