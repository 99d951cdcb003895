// iterate_wg.cuh -- the WARP-GROUP persistent kernel: eight warps work on one resident tile of trajectories.
//
// BASELINE.json north_star (1)/(3): "Riccati backward/forward sweeps run one CTA or warp-group per trajectory [group],
// with stage matrices staged in shared memory ... a persistent multi-iteration launch".  A CTA of eight warps owns a tile
// of TW <= 32 consecutive working-set columns (lane = problem, so every factor entry is ONE broadcast shared-memory read
// for the whole tile).  The tile's z, u, d rows live in shared memory for the whole launch (loaded once, stored once:
// no HBM traffic per iteration) together with a 6-row-per-stage scratch array g.
//
// What the eight warps buy.  One thread per problem runs an iteration as ONE dependent instruction stream (~11 k
// instructions; a lone warp needs 46 k cycles for it, iterate_res.cuh).  But only a small part of an iteration is
// inherently sequential over the stages: the recurrences  g_k = pre_k + Acl_k' g_{k+1}  (backward) and
// a_k = d_k + K_k s_k,  s_{k+1} = A_k s_k + B_k a_k  (forward).  Everything else is independent across stages and is
// split over the eight warps (warp w takes stages k = w mod 8):
//   P1  all warps   ra_k = z_k - u_k;  pre_k = K_k' ra_k  (the leading terms of g_k's accumulators);  dpre_k = Hinv_k ra_k
//   C1  warp 0 / 1  in-plane / cross-track recurrence  g_k = pre_k + Acl_k' g_{k+1}   (4 / 2 dependent FMAs per stage)
//   P2  all warps   d_k = dpre_k + E_k g_{k+1}
//   C2  warp 0 / 1  a_k = d_k + K_k s_k,  s_{k+1} = A_k s_k + B_k a_k   (the decoupled model makes the two warps independent)
//   P3  all warps   relaxation, prox, dual ascent of control block k (+ the terminal blocks); ds_k = z+ - z kept for C3
//   C3  warps 0-2   the five norm accumulators, each summed over the blocks in the oracle's order
//   then every warp evaluates the stopping test / rho update for its lane (identical arithmetic, no broadcast needed).
// Each accumulator still receives exactly the oracle's operations in the oracle's order -- P1 + C1 continue ONE fma
// chain per entry of g, P1 + P2 one per entry of d -- so the results are bit-identical to oracle/admm_ocp_cpu.c.
//
// Scope: shared decoupled factor (CW / Yamanaka-Ankersen structure, checked on the factor), "states unsplit, controls
// split" pattern, no affine term, no linear cost, shared parameter table: the pattern of BASELINE configs 1, 2, 3, 5.
#pragma once
#include "kernels.cuh"

namespace admmb {

constexpr int WG_WARPS = 8;                      // two per SM sub-partition: the stage-parallel phases interleave
constexpr int WG_MIN_TW = 8;                     // narrower tiles waste more than 3/4 of every warp: not worth it

struct WgLayout {                                // byte offsets inside the dynamic shared memory of one CTA
    int TW;                                      // tile width: problems per CTA
    int rz, rd, rg;                              // doubles per problem in the z / u, d and g arrays (odd: see below)
    size_t fac, par, typ, z, u, d, g, s0, nrm, total;
};

// Tile arrays are stored PROBLEM-major: the rows of one problem are contiguous, lane p starts at p * pitch.  A row is
// then a compile-time offset from a per-lane pointer (no index arithmetic on the sequential chains), and an odd pitch
// keeps a warp's 64-bit accesses conflict-free (16 lanes of a half-warp hit 16 distinct bank pairs).
__host__ __device__ inline WgLayout wg_layout(int N, int rows_zu, int TW)
{
    WgLayout L;
    L.TW = TW;
    L.rz = rows_zu | 1;
    L.rd = (3 * N) | 1;
    L.rg = (6 * N + 12) | 1;                     // 6 rows per stage + g_N / s_N (6) + terminal ds (6)
    const int nsb = rows_zu / 3;                 // split blocks: N controls + the split terminal blocks
    size_t o = 32;                               // [0, 8): mbarrier of the factor copy
    L.fac = o; o += sizeof(double) * FD * (size_t)N;
    L.par = o; o += sizeof(double) * 8 * (size_t)nsb;          // parameter table, compact: one record per split block
    L.typ = o; o += sizeof(int) * (size_t)((nsb + 3) / 4 * 4);
    const size_t col = sizeof(double) * (size_t)TW;
    L.z = o; o += col * L.rz;
    L.u = o; o += col * L.rz;
    L.d = o; o += col * L.rd;
    L.g = o; o += col * L.rg;
    L.s0 = o; o += col * 7;
    L.nrm = o; o += col * 5;
    L.total = (o + 15) / 16 * 16;
    return L;
}

// Shared-memory accesses by 32-bit shared-window address.  volatile + "memory": the compiler keeps them in program order
// (so the software pipelining written below is the order ptxas sees) and never treats a load as loop-invariant.
__device__ __forceinline__ double wg_ld(uint32_t a)
{
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void wg_st(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void wg_ld2(uint32_t a, double &x, double &y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void wg_ld4(uint32_t a, double (&r)[4])
{
    wg_ld2(a, r[0], r[1]);
    wg_ld2(a + 16, r[2], r[3]);
}

// Stage loop with the operands of the next stage loaded (into registers) BEFORE the current stage's arithmetic and
// stores: a warp issues in order, so without this every stage would start by waiting out its own load latency.
// Stage j = 0 .. count-1 is k = k0 + j * stride.  Branch-free body: the prefetch past the end re-reads the last stage.
template <class In, class LoadFn, class WorkFn>
__device__ __forceinline__ void wg_stage_loop(const int k0, const int stride, const int count, LoadFn load, WorkFn work)
{
    if (count <= 0) return;
    In a, b;
    load(k0, a);
    const int klast = k0 + (count - 1) * stride;
    int k = k0;
    for (int j = 0; j + 1 < count; j += 2, k += 2 * stride) {
        load(k + stride, b);
        work(k, a);
        load(j + 2 < count ? k + 2 * stride : klast, a);
        work(k + stride, b);
    }
    if (count & 1) work(klast, a);
}

template <bool ADAPT>
__global__ void __launch_bounds__(WG_WARPS * 32, 1)
k_admm_iterate_wg(const __grid_constant__ IterParams P, const int TW)
{
    extern __shared__ __align__(16) unsigned char wg_smem[];
    const int N = P.N;
    const WgLayout L = wg_layout(N, P.rows_zu, TW);
    const int tid = threadIdx.x, warp = tid >> 5;
    // lanes beyond the tile width are exact clones of the tile's last lane (same problem, same state, same values
    // written to the same addresses), so no access has to be predicated on the lane
    const int lane = min(tid & 31, TW - 1);
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(wg_smem);
    const uint32_t mbar = sm0;
    const uint32_t fac_s = sm0 + (uint32_t)L.fac, par_s = sm0 + (uint32_t)L.par, typ_s = sm0 + (uint32_t)L.typ;
    // this lane's problem in every tile array (byte addresses; row r of an array is at base + 8 r)
    const uint32_t zs = sm0 + (uint32_t)L.z + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t us = sm0 + (uint32_t)L.u + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t ds = sm0 + (uint32_t)L.d + (uint32_t)(lane * L.rd) * 8u;
    const uint32_t gs = sm0 + (uint32_t)L.g + (uint32_t)(lane * L.rg) * 8u;
    const uint32_t s0s = sm0 + (uint32_t)L.s0 + (uint32_t)(lane * 7) * 8u;
    const uint32_t nrs = sm0 + (uint32_t)L.nrm + (uint32_t)(lane * 5) * 8u;
    const uint32_t gN = gs + (uint32_t)(6 * N) * 8u;      // g_N, then s_N (6 rows), then the terminal blocks' ds (6 rows)

    // ---- once per CTA: factor (one TMA bulk copy), compact parameter table and block types
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_expect_tx(mbar, (uint32_t)(FD * N * 8));
        bulk_g2s(fac_s, P.fac_dec, (uint32_t)(FD * N * 8), mbar);
    }
    {
        double *parS = reinterpret_cast<double *>(wg_smem + L.par);
        int *typS = reinterpret_cast<int *>(wg_smem + L.typ);
        for (int b = tid; b < P.nb; b += WG_WARPS * 32) {
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

    const size_t ld = P.ld;
    const int ntiles = (P.n_active + TW - 1) / TW;
    const int nmine = (N - warp + WG_WARPS - 1) / WG_WARPS;             // stages k = warp, warp + 8, ... of the P phases
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TW + lane;
        const bool valid = t < P.n_active;
        const size_t p = (size_t)t;
        const bool live = valid && P.status[t] == ST_RUNNING;
        if (!__syncthreads_or(live)) continue;           // (also fences the previous tile's shared-memory traffic)

        // ---- tile in: the z, u, d rows of this tile's problems (row r by warp r mod 8; a warp reads TW consecutive
        //      doubles of one row; the odd per-problem pitch makes the transposing stores conflict-free)
        if (valid) {
#pragma unroll 4
            for (int r = warp; r < P.rows_zu; r += WG_WARPS) {
                wg_st(zs + 8u * r, __ldcg(P.z + (size_t)r * ld + p));
                wg_st(us + 8u * r, __ldcg(P.u + (size_t)r * ld + p));
            }
#pragma unroll 4
            for (int r = warp; r < 3 * N; r += WG_WARPS) wg_st(ds + 8u * r, __ldcg(P.d + (size_t)r * ld + p));
            if (warp == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) wg_st(s0s + 8u * i, P.s0[(size_t)i * ld + p]);
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
            if (!__any_sync(0xffffffffu, run)) break;    // every warp holds the same flags: the exit is uniform
            const double rinv = 1.0 / rho;

            // ================= P1: stage-parallel head of the backward sweep
            {
                struct In { double z[3], u[3], k0[4], k1[4], kc[2], h0[2], h1[2], hc[2]; };
                wg_stage_loop<In>(warp, WG_WARPS, nmine,
                    [&](int k, In &in) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                        wg_ld4(fk + D_KIN * 8, in.k0);
                        wg_ld4(fk + (D_KIN + 4) * 8, in.k1);
                        wg_ld2(fk + D_KC * 8, in.kc[0], in.kc[1]);
                        wg_ld2(fk + D_HIN * 8, in.h0[0], in.h0[1]);
                        wg_ld2(fk + (D_HIN + 2) * 8, in.h1[0], in.h1[1]);
                        wg_ld2(fk + D_HC * 8, in.hc[0], in.hc[1]);
                        const uint32_t o = (uint32_t)(3 * k) * 8u;
#pragma unroll
                        for (int e = 0; e < 3; ++e) { in.z[e] = wg_ld(zs + o + 8u * e); in.u[e] = wg_ld(us + o + 8u * e); }
                    },
                    [&](int k, const In &in) {
                        double ra[3];
#pragma unroll
                        for (int e = 0; e < 3; ++e) {
                            const double uu = ADAPT ? in.u[e] * sigma : in.u[e];
                            ra[e] = in.z[e] - uu;
                        }
                        double d0 = in.h0[0] * ra[0];
                        d0 = fma(in.h0[1], ra[1], d0);
                        double d1 = in.h1[0] * ra[0];
                        d1 = fma(in.h1[1], ra[1], d1);
                        const double d2 = in.hc[0] * ra[2];
                        const uint32_t dk = ds + (uint32_t)(3 * k) * 8u, gk = gs + (uint32_t)(6 * k) * 8u;
                        if (run) { wg_st(dk, d0); wg_st(dk + 8, d1); wg_st(dk + 16, d2); }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            double pi = fma(in.k0[i], ra[0], 0.0);
                            pi = fma(in.k1[i], ra[1], pi);
                            wg_st(gk + 8u * i, pi);
                        }
                        wg_st(gk + 32, fma(in.kc[0], ra[2], 0.0));
                        wg_st(gk + 40, fma(in.kc[1], ra[2], 0.0));
                    });
            }
            if (warp == WG_WARPS - 1) {                  // g_N: right-hand side of the terminal blocks
                double tv[6];
#pragma unroll
                for (int tb = 0; tb < 2; ++tb) {
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) != BLK_NONE) {
                        const uint32_t o = (uint32_t)((de >> 8) * 3) * 8u;
#pragma unroll
                        for (int e = 0; e < 3; ++e) {
                            double uu = wg_ld(us + o + 8u * e);
                            if (ADAPT) uu = uu * sigma;
                            tv[3 * tb + e] = wg_ld(zs + o + 8u * e) - uu;
                        }
                    } else {
                        tv[3 * tb] = 0.0; tv[3 * tb + 1] = 0.0; tv[3 * tb + 2] = 0.0;
                    }
                }
                wg_st(gN, tv[0]); wg_st(gN + 8, tv[1]); wg_st(gN + 16, tv[3]);          // in-plane: s0 s1 s3 s4
                wg_st(gN + 24, tv[4]); wg_st(gN + 32, tv[2]); wg_st(gN + 40, tv[5]);     // cross-track: s2 s5
            }
            __syncthreads();

            // ================= C1: the backward recurrence (in-plane on warp 0, cross-track on warp 1)
            if (warp == 0) {
                double gi[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) gi[i] = wg_ld(gN + 8u * i);
                struct In { double pr[4], ac[4][4]; };
                wg_stage_loop<In>(N - 1, -1, N,
                    [&](int k, In &in) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD + D_ACLIN) * 8u, gk = gs + (uint32_t)(6 * k) * 8u;
#pragma unroll
                        for (int l = 0; l < 4; ++l) wg_ld4(fk + 32u * l, in.ac[l]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) in.pr[i] = wg_ld(gk + 8u * i);
                    },
                    [&](int k, const In &in) {
                        double pi[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) pi[i] = in.pr[i];
#pragma unroll
                        for (int l = 0; l < 4; ++l)
#pragma unroll
                            for (int i = 0; i < 4; ++i) pi[i] = fma(in.ac[l][i], gi[l], pi[i]);
                        const uint32_t gk = gs + (uint32_t)(6 * k) * 8u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) { gi[i] = pi[i]; wg_st(gk + 8u * i, pi[i]); }
                    });
            } else if (warp == 1) {
                double gc0 = wg_ld(gN + 32), gc1 = wg_ld(gN + 40);
                struct In { double a0[2], a1[2], p[2]; };
                wg_stage_loop<In>(N - 1, -1, N,
                    [&](int k, In &in) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD + D_ACLC) * 8u, gk = gs + (uint32_t)(6 * k) * 8u;
                        wg_ld2(fk, in.a0[0], in.a0[1]);
                        wg_ld2(fk + 16, in.a1[0], in.a1[1]);
                        in.p[0] = wg_ld(gk + 32);
                        in.p[1] = wg_ld(gk + 40);
                    },
                    [&](int k, const In &in) {
                        double p0 = fma(in.a0[0], gc0, in.p[0]), p1 = fma(in.a0[1], gc0, in.p[1]);
                        p0 = fma(in.a1[0], gc1, p0); p1 = fma(in.a1[1], gc1, p1);
                        gc0 = p0; gc1 = p1;
                        const uint32_t gk = gs + (uint32_t)(6 * k) * 8u;
                        wg_st(gk + 32, p0); wg_st(gk + 40, p1);
                    });
            }
            __syncthreads();

            // ================= P2: d_k = dpre_k + E_k g_{k+1}
            {
                struct In { double e0[4], e1[4], ec[2], g[6], d[3]; };
                wg_stage_loop<In>(warp, WG_WARPS, nmine,
                    [&](int k, In &in) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                        wg_ld4(fk + D_EIN * 8, in.e0);
                        wg_ld4(fk + (D_EIN + 4) * 8, in.e1);
                        wg_ld2(fk + D_EC * 8, in.ec[0], in.ec[1]);
                        const uint32_t gk = gs + (uint32_t)(6 * (k + 1)) * 8u, dk = ds + (uint32_t)(3 * k) * 8u;
#pragma unroll
                        for (int i = 0; i < 6; ++i) in.g[i] = wg_ld(gk + 8u * i);
#pragma unroll
                        for (int e = 0; e < 3; ++e) in.d[e] = wg_ld(dk + 8u * e);
                    },
                    [&](int k, const In &in) {
                        double d0 = in.d[0], d1 = in.d[1], d2 = in.d[2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { d0 = fma(in.e0[i], in.g[i], d0); d1 = fma(in.e1[i], in.g[i], d1); }
                        d2 = fma(in.ec[0], in.g[4], d2);
                        d2 = fma(in.ec[1], in.g[5], d2);
                        const uint32_t dk = ds + (uint32_t)(3 * k) * 8u;
                        if (run) { wg_st(dk, d0); wg_st(dk + 8, d1); wg_st(dk + 16, d2); }
                    });
            }
            __syncthreads();

            // ================= C2: the forward recurrence (in-plane on warp 0, cross-track on warp 1)
            if (warp == 0) {
                double si[4] = {wg_ld(s0s), wg_ld(s0s + 8), wg_ld(s0s + 24), wg_ld(s0s + 32)};
                struct In { double d[2], kn[2][4], an[4][4], bn[4][2]; };
                wg_stage_loop<In>(0, 1, N,
                    [&](int k, In &o) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u, dk = ds + (uint32_t)(3 * k) * 8u;
                        wg_ld4(fk + D_KIN * 8, o.kn[0]);
                        wg_ld4(fk + (D_KIN + 4) * 8, o.kn[1]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            wg_ld4(fk + (D_AIN + 4 * r) * 8, o.an[r]);
                            wg_ld2(fk + (D_BIN + 2 * r) * 8, o.bn[r][0], o.bn[r][1]);
                        }
                        o.d[0] = wg_ld(dk);
                        o.d[1] = wg_ld(dk + 8);
                    },
                    [&](int k, const In &o) {
                        double a0 = o.d[0], a1 = o.d[1];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { a0 = fma(o.kn[0][i], si[i], a0); a1 = fma(o.kn[1][i], si[i], a1); }
                        const uint32_t gk = gs + (uint32_t)(6 * k) * 8u;
                        wg_st(gk, a0);
                        wg_st(gk + 8, a1);
                        double ni[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            double acc = o.an[r][0] * si[0];
                            acc = fma(o.an[r][1], si[1], acc);
                            acc = fma(o.an[r][2], si[2], acc);
                            acc = fma(o.an[r][3], si[3], acc);
                            acc = fma(o.bn[r][0], a0, acc);
                            acc = fma(o.bn[r][1], a1, acc);
                            ni[r] = acc;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) si[r] = ni[r];
                    });
                wg_st(gN, si[0]); wg_st(gN + 8, si[1]); wg_st(gN + 24, si[2]); wg_st(gN + 32, si[3]);   // s_N, natural order
            } else if (warp == 1) {
                double sc0 = wg_ld(s0s + 16), sc1 = wg_ld(s0s + 40);
                struct In { double kc[2], a0[2], a1[2], bc[2], d2; };
                wg_stage_loop<In>(0, 1, N,
                    [&](int k, In &in) {
                        const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                        wg_ld2(fk + D_KC * 8, in.kc[0], in.kc[1]);
                        wg_ld2(fk + D_AC * 8, in.a0[0], in.a0[1]);
                        wg_ld2(fk + (D_AC + 2) * 8, in.a1[0], in.a1[1]);
                        wg_ld2(fk + D_BC * 8, in.bc[0], in.bc[1]);
                        in.d2 = wg_ld(ds + (uint32_t)(3 * k + 2) * 8u);
                    },
                    [&](int k, const In &in) {
                        double a2 = fma(in.kc[0], sc0, in.d2);
                        a2 = fma(in.kc[1], sc1, a2);
                        wg_st(gs + (uint32_t)(6 * k + 2) * 8u, a2);
                        double n0 = in.a0[0] * sc0;
                        n0 = fma(in.a0[1], sc1, n0);
                        n0 = fma(in.bc[0], a2, n0);
                        double n1 = in.a1[0] * sc0;
                        n1 = fma(in.a1[1], sc1, n1);
                        n1 = fma(in.bc[1], a2, n1);
                        sc0 = n0; sc1 = n1;
                    });
                wg_st(gN + 16, sc0); wg_st(gN + 40, sc1);
            }
            __syncthreads();

            // ================= P3: relaxation, prox, dual ascent per block; ds = z+ - z kept for the s-norm
            {
                struct In { double x[3], z[3], u[3], pr[8]; int type; };
                // block `slot` (compact index) with x at xa, ds written to da
                auto load_block = [&](int slot, uint32_t xa, In &in) {
                    const uint32_t pa = par_s + (uint32_t)slot * 64u, o = (uint32_t)(3 * slot) * 8u;
                    wg_ld2(pa, in.pr[0], in.pr[1]); wg_ld2(pa + 16, in.pr[2], in.pr[3]);
                    wg_ld2(pa + 32, in.pr[4], in.pr[5]); wg_ld2(pa + 48, in.pr[6], in.pr[7]);
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(in.type) : "r"(typ_s + 4u * slot) : "memory");
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        in.x[e] = wg_ld(xa + 8u * e);
                        in.z[e] = wg_ld(zs + o + 8u * e);
                        in.u[e] = wg_ld(us + o + 8u * e);
                    }
                };
                auto update_block = [&](int slot, uint32_t da, const In &in) {
                    double v[3], zn[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double uo = ADAPT ? in.u[e] * sigma : in.u[e];
                        const double xh = fma(P.alpha, in.x[e], P.oma * in.z[e]);
                        v[e] = xh + uo;
                    }
                    prox_block_dev(in.type, [&](int q) { return in.pr[q]; }, rinv, v, zn);
                    const uint32_t o = (uint32_t)(3 * slot) * 8u;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        wg_st(da + 8u * e, zn[e] - in.z[e]);
                        if (run) { wg_st(zs + o + 8u * e, zn[e]); wg_st(us + o + 8u * e, v[e] - zn[e]); }
                    }
                };
                // control block k has compact index k (fast pattern); its x = a_k is at g rows 6k..6k+2, ds goes to 6k+3..
                wg_stage_loop<In>(warp, WG_WARPS, nmine,
                    [&](int k, In &in) { load_block(k, gs + (uint32_t)(6 * k) * 8u, in); },
                    [&](int k, const In &in) { update_block(k, gs + (uint32_t)(6 * k + 3) * 8u, in); });
                if (warp >= WG_WARPS - 2) {                  // the two terminal blocks: x = s_N
                    const int tb = warp - (WG_WARPS - 2);
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) != BLK_NONE) {
                        In in;
                        load_block(de >> 8, gN + (uint32_t)(3 * tb) * 8u, in);
                        update_block(de >> 8, gN + (uint32_t)(6 + 3 * tb) * 8u, in);
                    }
                }
            }
            __syncthreads();

            // ================= C3: the five norm accumulators, blocks in the oracle's order
            if (warp < 3) {
                double acc0 = 0.0, acc1 = 0.0;
                auto add_block = [&](uint32_t xa, uint32_t za, uint32_t ua, uint32_t da) {
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (warp == 0) {
                            const double dr = wg_ld(xa + 8u * e) - wg_ld(za + 8u * e), dd = wg_ld(da + 8u * e);
                            acc0 = fma(dr, dr, acc0);
                            acc1 = fma(dd, dd, acc1);
                        } else if (warp == 1) {
                            const double x = wg_ld(xa + 8u * e), z = wg_ld(za + 8u * e);
                            acc0 = fma(x, x, acc0);
                            acc1 = fma(z, z, acc1);
                        } else {
                            const double un = wg_ld(ua + 8u * e);
                            acc0 = fma(un, un, acc0);
                        }
                    }
                };
#pragma unroll 4
                for (int k = 0; k < N; ++k)
                    add_block(gs + (uint32_t)(6 * k) * 8u, zs + (uint32_t)(3 * k) * 8u, us + (uint32_t)(3 * k) * 8u,
                              gs + (uint32_t)(6 * k + 3) * 8u);
#pragma unroll
                for (int tb = 0; tb < 2; ++tb) {
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) == BLK_NONE) continue;
                    const uint32_t o = (uint32_t)(3 * (de >> 8)) * 8u;
                    add_block(gN + (uint32_t)(3 * tb) * 8u, zs + o, us + o, gN + (uint32_t)(6 + 3 * tb) * 8u);
                }
                if (warp == 0) { wg_st(nrs, acc0); wg_st(nrs + 8, acc1); }
                else if (warp == 1) { wg_st(nrs + 16, acc0); wg_st(nrs + 24, acc1); }
                else wg_st(nrs + 32, acc0);
            }
            __syncthreads();

            // ================= stopping test and rho update: every warp, for its own copy of the lane's state
            if (run) {
                ++it;
                sigma = 1.0;
                r_norm = sqrt(wg_ld(nrs));
                s_norm = rho * sqrt(wg_ld(nrs + 8));
                const double nx = sqrt(wg_ld(nrs + 16)), nz = sqrt(wg_ld(nrs + 24));
                eps_pri = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
                eps_dual = fma(P.reltol, rho * sqrt(wg_ld(nrs + 32)), P.sqrtn_abs);
                if (P.hist && warp == 0) {
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
                    // shared factor: independent of rho here (P = 0), so a rho change needs no refactorisation
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
        }

        // ---- tile out (the loop ends behind a barrier or before any phase of a new iteration has started)
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int r = warp; r < P.rows_zu; r += WG_WARPS) {
                __stcg(P.z + (size_t)r * ld + p, wg_ld(zs + 8u * r));
                __stcg(P.u + (size_t)r * ld + p, wg_ld(us + 8u * r));
            }
#pragma unroll 4
            for (int r = warp; r < 3 * N; r += WG_WARPS) __stcg(P.d + (size_t)r * ld + p, wg_ld(ds + 8u * r));
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
