// iterate_wg.cuh -- the WARP-GROUP persistent kernel: eight role-split warps work on one resident tile of trajectories.
//
// BASELINE.json north_star (1)/(3): "Riccati backward/forward sweeps run one CTA or warp-group per trajectory [group],
// with stage matrices staged in shared memory ... a persistent multi-iteration launch".  A CTA of eight warps owns a tile
// of TW <= 32 consecutive working-set columns (lane = problem, so every factor entry is ONE broadcast shared-memory read
// for the whole tile).  The tile's z, u, d rows live in shared memory for the whole launch (loaded once, stored once:
// no HBM traffic per iteration) together with a 6-row-per-stage scratch array g.
//
// What bounds an iteration is the length of its sequential part, so the roles are cut along the dependency chains:
//   warp 0  CHAIN, in-plane     backward sweep (ra, d, g recurrence: 36 FMAs per stage) then, without any barrier, the
//   warp 1  CHAIN, cross-track  forward sweep (a_k = d_k + K_k s_k, s_{k+1} = A_k s_k + B_k a_k).  The decoupled model makes
//                               the two warps independent of each other; each reads back only the d rows it wrote itself.
//   warps 2, 3, 6, 7  PROX      trail the forward sweep stage by stage: relaxation, prox and dual ascent of control block k
//                               as soon as both chain warps have published a_k (one mbarrier per stage), then the terminal
//                               blocks; z+ - z is left in g for the norms.
//   warp 5  NORMS               trails the prox warps block by block: the five norm accumulators, each summed over the
//                               blocks in the oracle's order.
//   warp 4                      shares its sub-partition with the in-plane chain and stays out of its way (tile I/O only).
// One CTA-wide barrier per iteration (before the stopping test); every warp then evaluates the test for its lane
// (identical arithmetic, no broadcast needed).  The per-stage hand-over uses mbarriers (arrive.release by one elected lane
// after __syncwarp, try_wait.acquire by the consumers), whose phase parity flips once per iteration.
// Each accumulator still receives exactly the oracle's operations in the oracle's order, so the results are bit-identical
// to oracle/admm_ocp_cpu.c.
//
// Scope: shared decoupled factor (CW / Yamanaka-Ankersen structure, checked on the factor), "states unsplit, controls
// split" pattern, no affine term, no linear cost, shared parameter table: the pattern of BASELINE configs 1, 2, 3, 5.
#pragma once
#include "kernels.cuh"

namespace admmb {

constexpr int WG_WARPS = 8;
constexpr int WG_MIN_TW = 8;                     // narrower tiles waste more than 3/4 of every warp: not worth it
constexpr int WG_PROX = 4;                       // prox warps

struct WgLayout {                                // byte offsets inside the dynamic shared memory of one CTA
    int TW;                                      // tile width: problems per CTA
    int rz, rd, rg;                              // doubles per problem in the z / u, d and g arrays (odd: see below)
    size_t bars, fac, par, typ, z, u, d, g, s0, nrm, total;
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
    L.rg = (6 * N + 12) | 1;                     // per stage a_k (3) + ds_k (3); then s_N (6) and the terminal ds (6)
    const int nsb = rows_zu / 3;                 // split blocks: N controls + the split terminal blocks
    size_t o = 16;                               // [0, 8): mbarrier of the factor copy
    L.bars = o; o += 8 * (size_t)(2 * N + 2);    // bC2[0..N] (a_k / s_N published), bP3[0..N] (block k / terminal blocks done)
    o = (o + 15) / 16 * 16;
    L.fac = o; o += sizeof(double) * FD * (size_t)N;
    L.par = o; o += sizeof(double) * 8 * (size_t)nsb;          // parameter table, compact: one record per split block
    L.typ = o; o += sizeof(int) * (size_t)((nsb + 3) / 4 * 4);
    const size_t col = sizeof(double) * (size_t)TW;
    L.z = o; o += col * L.rz;
    L.u = o; o += col * L.rz;
    L.d = o; o += col * L.rd;
    L.g = o; o += col * L.rg;
    L.s0 = o; o += col * 7;
    L.nrm = o; o += col * 10;                    // two sets of five, alternating by iteration parity
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

// Per-stage hand-over between warps.  The producer warp has written shared memory with all its lanes; __syncwarp orders
// those writes before the elected lane's arrive (release, CTA scope), the consumers' try_wait (acquire) orders their
// reads after it.  A wait that never completes is a bug: it traps instead of hanging the GPU.
__device__ __forceinline__ void wg_arrive(uint32_t bar)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1 << 22)) __trap();
    } while (!ok);
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
    const uint32_t bC2 = sm0 + (uint32_t)L.bars, bP3 = bC2 + 8u * (uint32_t)(N + 1);
    const uint32_t fac_s = sm0 + (uint32_t)L.fac, par_s = sm0 + (uint32_t)L.par, typ_s = sm0 + (uint32_t)L.typ;
    // this lane's problem in every tile array (byte addresses; row r of an array is at base + 8 r)
    const uint32_t zs = sm0 + (uint32_t)L.z + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t us = sm0 + (uint32_t)L.u + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t ds = sm0 + (uint32_t)L.d + (uint32_t)(lane * L.rd) * 8u;
    const uint32_t gs = sm0 + (uint32_t)L.g + (uint32_t)(lane * L.rg) * 8u;
    const uint32_t s0s = sm0 + (uint32_t)L.s0 + (uint32_t)(lane * 7) * 8u;
    const uint32_t nrs0 = sm0 + (uint32_t)L.nrm + (uint32_t)(lane * 10) * 8u;
    const uint32_t gN = gs + (uint32_t)(6 * N) * 8u;      // s_N in natural order (6 rows), then the terminal blocks' ds (6 rows)

    // roles
    const bool chain_in = warp == 0, chain_c = warp == 1, norms = warp == 5;
    const int prox = warp == 2 ? 0 : warp == 3 ? 1 : warp == 6 ? 2 : warp == 7 ? 3 : -1;

    // ---- once per CTA: factor (one TMA bulk copy), hand-over barriers, compact parameter table and block types
    if (tid == 0) {
        mbar_init(mbar, 1);
        for (int k = 0; k <= N; ++k) {
            mbar_init(bC2 + 8u * k, 2);                       // both chain warps
            mbar_init(bP3 + 8u * k, k < N ? 1 : 2);           // the prox warp of block k; the two terminal blocks
        }
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
    uint32_t ph = 0;                                     // parity of the hand-over barriers: flips once per iteration
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
            // norm slots of this iteration (the other set may still be read by a warp that is late in the previous
            // iteration's stopping test)
            const uint32_t nrs = nrs0 + 40u * ph;

            if (chain_in) {
                // ================= in-plane chain: states (s0, s1, s3, s4), controls (a0, a1)
                // ---- backward sweep: ra_k = z_k - u_k; d_k = Hinv_k ra_k + E_k g_{k+1}; g_k = K_k' ra_k + Acl_k' g_{k+1}
                double gi[4];
                {
                    double tv[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int tb = 0; tb < 2; ++tb) {
                        const int de = tb == 0 ? de_t0 : de_t1;
                        if ((de & 0xff) != BLK_NONE) {
                            const uint32_t o = (uint32_t)((de >> 8) * 3) * 8u;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                double uu = wg_ld(us + o + 8u * e);
                                if (ADAPT) uu = uu * sigma;
                                tv[2 * tb + e] = wg_ld(zs + o + 8u * e) - uu;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) gi[i] = tv[i];
                }
                {
                    struct In { double z[2], u[2], h0[2], h1[2], e0[4], e1[4], k0[4], k1[4], ac[4][4]; };
                    wg_stage_loop<In>(N - 1, -1, N,
                        [&](int k, In &in) {
                            const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                            wg_ld4(fk + D_KIN * 8, in.k0);
                            wg_ld4(fk + (D_KIN + 4) * 8, in.k1);
#pragma unroll
                            for (int l = 0; l < 4; ++l) wg_ld4(fk + (D_ACLIN + 4 * l) * 8, in.ac[l]);
                            wg_ld2(fk + D_HIN * 8, in.h0[0], in.h0[1]);
                            wg_ld2(fk + (D_HIN + 2) * 8, in.h1[0], in.h1[1]);
                            wg_ld4(fk + D_EIN * 8, in.e0);
                            wg_ld4(fk + (D_EIN + 4) * 8, in.e1);
                            const uint32_t o = (uint32_t)(3 * k) * 8u;
#pragma unroll
                            for (int e = 0; e < 2; ++e) { in.z[e] = wg_ld(zs + o + 8u * e); in.u[e] = wg_ld(us + o + 8u * e); }
                        },
                        [&](int k, const In &in) {
                            double ra[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const double uu = ADAPT ? in.u[e] * sigma : in.u[e];
                                ra[e] = in.z[e] - uu;
                            }
                            double pi[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                pi[i] = fma(in.k0[i], ra[0], 0.0);
                                pi[i] = fma(in.k1[i], ra[1], pi[i]);
                            }
#pragma unroll
                            for (int l = 0; l < 4; ++l)
#pragma unroll
                                for (int i = 0; i < 4; ++i) pi[i] = fma(in.ac[l][i], gi[l], pi[i]);
                            double d0 = in.h0[0] * ra[0];
                            d0 = fma(in.h0[1], ra[1], d0);
                            double d1 = in.h1[0] * ra[0];
                            d1 = fma(in.h1[1], ra[1], d1);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { d0 = fma(in.e0[i], gi[i], d0); d1 = fma(in.e1[i], gi[i], d1); }
                            const uint32_t dk = ds + (uint32_t)(3 * k) * 8u;
                            if (run) { wg_st(dk, d0); wg_st(dk + 8, d1); }
#pragma unroll
                            for (int i = 0; i < 4; ++i) gi[i] = pi[i];
                        });
                }
                // ---- forward sweep: a_k = d_k + K_k s_k (published for the prox warps), s_{k+1} = A_k s_k + B_k a_k
                double si[4] = {wg_ld(s0s), wg_ld(s0s + 8), wg_ld(s0s + 24), wg_ld(s0s + 32)};
                {
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
                            wg_arrive(bC2 + 8u * k);
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
                }
                wg_st(gN, si[0]); wg_st(gN + 8, si[1]); wg_st(gN + 24, si[2]); wg_st(gN + 32, si[3]);   // s_N, natural order
                wg_arrive(bC2 + 8u * N);
            } else if (chain_c) {
                // ================= cross-track chain: states (s2, s5), control a2
                double gc0 = 0.0, gc1 = 0.0;
#pragma unroll
                for (int tb = 0; tb < 2; ++tb) {
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) != BLK_NONE) {
                        const uint32_t o = (uint32_t)((de >> 8) * 3 + 2) * 8u;
                        double uu = wg_ld(us + o);
                        if (ADAPT) uu = uu * sigma;
                        const double v = wg_ld(zs + o) - uu;
                        if (tb == 0) gc0 = v; else gc1 = v;
                    }
                }
                {
                    struct In { double z, u, hc[2], ec[2], kc[2], a0[2], a1[2]; };
                    wg_stage_loop<In>(N - 1, -1, N,
                        [&](int k, In &in) {
                            const uint32_t fk = fac_s + (uint32_t)(k * FD) * 8u;
                            wg_ld2(fk + D_KC * 8, in.kc[0], in.kc[1]);
                            wg_ld2(fk + D_ACLC * 8, in.a0[0], in.a0[1]);
                            wg_ld2(fk + (D_ACLC + 2) * 8, in.a1[0], in.a1[1]);
                            wg_ld2(fk + D_HC * 8, in.hc[0], in.hc[1]);
                            wg_ld2(fk + D_EC * 8, in.ec[0], in.ec[1]);
                            const uint32_t o = (uint32_t)(3 * k + 2) * 8u;
                            in.z = wg_ld(zs + o);
                            in.u = wg_ld(us + o);
                        },
                        [&](int k, const In &in) {
                            const double uu = ADAPT ? in.u * sigma : in.u;
                            const double ra = in.z - uu;
                            double p0 = fma(in.kc[0], ra, 0.0), p1 = fma(in.kc[1], ra, 0.0);
                            p0 = fma(in.a0[0], gc0, p0); p1 = fma(in.a0[1], gc0, p1);
                            p0 = fma(in.a1[0], gc1, p0); p1 = fma(in.a1[1], gc1, p1);
                            double d2 = in.hc[0] * ra;
                            d2 = fma(in.ec[0], gc0, d2);
                            d2 = fma(in.ec[1], gc1, d2);
                            if (run) wg_st(ds + (uint32_t)(3 * k + 2) * 8u, d2);
                            gc0 = p0; gc1 = p1;
                        });
                }
                double sc0 = wg_ld(s0s + 16), sc1 = wg_ld(s0s + 40);
                {
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
                            wg_arrive(bC2 + 8u * k);
                            double n0 = in.a0[0] * sc0;
                            n0 = fma(in.a0[1], sc1, n0);
                            n0 = fma(in.bc[0], a2, n0);
                            double n1 = in.a1[0] * sc0;
                            n1 = fma(in.a1[1], sc1, n1);
                            n1 = fma(in.bc[1], a2, n1);
                            sc0 = n0; sc1 = n1;
                        });
                }
                wg_st(gN + 16, sc0); wg_st(gN + 40, sc1);
                wg_arrive(bC2 + 8u * N);
            } else if (prox >= 0) {
                // ================= prox warps: relaxation, prox, dual ascent of block k behind the forward sweep
                struct In { double x[3], z[3], u[3], pr[8]; int type; };
                // block `slot` (compact index) with x at xa; ds = z+ - z goes to da
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
                for (int k = prox; k < N; k += WG_PROX) {
                    In in;
                    wg_wait(bC2 + 8u * k, ph);
                    load_block(k, gs + (uint32_t)(6 * k) * 8u, in);
                    update_block(k, gs + (uint32_t)(6 * k + 3) * 8u, in);
                    wg_arrive(bP3 + 8u * k);
                }
                if (prox < 2) {                              // the two terminal blocks: x = s_N
                    const int de = prox == 0 ? de_t0 : de_t1;
                    wg_wait(bC2 + 8u * N, ph);
                    if ((de & 0xff) != BLK_NONE) {
                        In in;
                        load_block(de >> 8, gN + (uint32_t)(3 * prox) * 8u, in);
                        update_block(de >> 8, gN + (uint32_t)(6 + 3 * prox) * 8u, in);
                    }
                    wg_arrive(bP3 + 8u * N);
                }
            } else if (norms) {
                // ================= the five norm accumulators, blocks in the oracle's order, behind the prox warps
                double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
                auto add_block = [&](uint32_t xa, uint32_t za, uint32_t ua, uint32_t da) {
                    double x[3], z[3], u[3], dd[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        x[e] = wg_ld(xa + 8u * e); z[e] = wg_ld(za + 8u * e);
                        u[e] = wg_ld(ua + 8u * e); dd[e] = wg_ld(da + 8u * e);
                    }
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double dr = x[e] - z[e];
                        rr = fma(dr, dr, rr);
                        ss = fma(dd[e], dd[e], ss);
                        xx = fma(x[e], x[e], xx);
                        zz = fma(z[e], z[e], zz);
                        uu = fma(u[e], u[e], uu);
                    }
                };
                for (int k = 0; k < N; ++k) {
                    wg_wait(bP3 + 8u * k, ph);
                    add_block(gs + (uint32_t)(6 * k) * 8u, zs + (uint32_t)(3 * k) * 8u, us + (uint32_t)(3 * k) * 8u,
                              gs + (uint32_t)(6 * k + 3) * 8u);
                }
                wg_wait(bP3 + 8u * N, ph);
#pragma unroll
                for (int tb = 0; tb < 2; ++tb) {
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) == BLK_NONE) continue;
                    const uint32_t o = (uint32_t)(3 * (de >> 8)) * 8u;
                    add_block(gN + (uint32_t)(3 * tb) * 8u, zs + o, us + o, gN + (uint32_t)(6 + 3 * tb) * 8u);
                }
                wg_st(nrs, rr); wg_st(nrs + 8, ss); wg_st(nrs + 16, xx); wg_st(nrs + 24, zz); wg_st(nrs + 32, uu);
            }
            ph ^= 1u;
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
