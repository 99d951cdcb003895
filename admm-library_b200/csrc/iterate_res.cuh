// iterate_res.cuh -- the RESIDENT-TILE persistent kernel: the iterates stay on chip for a whole launch.
//
// One warp owns a tile of 32 consecutive working-set columns (lane = problem).  The tile's z, u and d rows
// ((2 * rows_zu + 3N) x 32 doubles; 118 KB at N = 50) are brought into shared memory ONCE per launch by TMA
// (cp.async.bulk.tensor.2d boxes [rows] x [32 columns], SASS UTMALDG), every one of the launch's `chunk` iterations
// then reads and writes shared memory only (conflict-free LDS.64 / STS.64: lane p owns bank pair 2p), and the tile goes
// back with TMA stores (UTMASTG) when the launch ends or every problem of the tile has finished.  Per-iteration HBM
// traffic is zero; the bytes that move are 2 x 3.7 KB per problem per LAUNCH instead of 5 KB (8.3 KB measured, with
// the d round trip) per ITERATION.  With global-memory latency off the sequential stage chain a lone warp runs an
// iteration in a fraction of the time of k_admm_iterate's thread, which is what bounds narrow working sets
// (<= 16 k problems, i.e. the whole of a strong-scaled 65,536-problem solve on 8 GPUs).
//
// The arithmetic is admm_iteration_dec / admm_iteration_fast itself (kernels.cuh) instantiated with SmemIO, so every
// accumulator sees the oracle's operations in the oracle's order: results are bit-identical.
// Scope: shared factor staged in shared memory, "states unsplit / controls split" pattern (all five benchmark
// configurations); anything else keeps k_admm_iterate.  A finished problem simply stops storing (its columns in the
// tile are final); idle lanes cost nothing here because no global sector is ever partially written.
#pragma once
#include <cuda.h>
#include "kernels.cuh"

namespace admmb {

struct ResMaps {
    CUtensorMap z, u, d;          // [rows][ld] working-set arrays, boxes [box rows] x [32 columns]
    int zu_box, zu_nbox;          // rows_zu = zu_box * zu_nbox (exact boxes: no out-of-bounds rows)
    int d_box, d_nbox;            // 3N = d_box * d_nbox
};

struct SmemIO {
    double *zs, *us, *ds;         // the tile's rows, already offset by the lane
    const double *s0s;            // [6][32] copy of the tile's initial states, offset by the lane
    __device__ __forceinline__ double *z(const IterParams &, size_t) const { return zs; }
    __device__ __forceinline__ double *u(const IterParams &, size_t) const { return us; }
    __device__ __forceinline__ double *d(const IterParams &, size_t) const { return ds; }
    __device__ __forceinline__ size_t pitch(const IterParams &) const { return 32; }
    __device__ __forceinline__ double s0(const IterParams &, size_t, int i) const { return ld(s0s + 32 * i); }
    // plain accesses: the pointers derive from the kernel's shared array, so they compile to LDS.64 / STS.64, and the
    // compiler sees the loads and stores as memory operations (an asm load without a memory operand would be hoisted
    // out of the iteration loop as loop-invariant)
    static __device__ __forceinline__ double ld(const double *a) { return *a; }
    static __device__ __forceinline__ void st(double *a, double v) { *a = v; }
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int col, int row, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int col, int row, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(col), "r"(row), "r"(src) : "memory");
}

// bytes of dynamic shared memory: [16 B factor mbarrier][16 B tile mbarrier][factor][par][bdesc] pad 128 [z][u][d][s0]
__host__ __device__ inline size_t res_head_bytes(int N, int nb, bool decoupled, bool par_batched)
{
    size_t b = 32 + sizeof(double) * (size_t)(decoupled ? FD : FS) * N + (par_batched ? 0 : sizeof(double) * 8 * nb) +
               sizeof(int) * (size_t)((nb + 3) / 4) * 4;
    return (b + 127) / 128 * 128;
}
__host__ __device__ inline size_t res_smem_bytes(int N, int nb, int rows_zu, bool decoupled, bool par_batched)
{
    return res_head_bytes(N, nb, decoupled, par_batched) + (size_t)(2 * rows_zu + 3 * N + 6) * 256;
}

// MODE 1: coupled model (admm_iteration_fast), MODE 2: decoupled packed records (admm_iteration_dec)
template <bool HAS_C, bool HAS_Q, bool ADAPT, int MODE>
__global__ void __launch_bounds__(32, 1)
k_admm_iterate_res(const __grid_constant__ IterParams P, const __grid_constant__ ResMaps M)
{
    extern __shared__ __align__(128) unsigned char res_smem[];
    constexpr int REC = (MODE == 2) ? FD : FS;
    double *facS = reinterpret_cast<double *>(res_smem + 32);
    const double *fac_src = (MODE == 2) ? P.fac_dec : P.fac;
    double *parS = facS + (size_t)REC * P.N;
    int *bdS = reinterpret_cast<int *>(parS + (P.par_batched ? 0 : 8 * P.nb));
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(res_smem);
    const uint32_t tbar = mbar + 16;
    const uint32_t fac_sbase = (uint32_t)__cvta_generic_to_shared(facS);
    const uint32_t par_sbase = (uint32_t)__cvta_generic_to_shared(parS);
    const uint32_t bd_bytes = (uint32_t)(((P.nb + 3) / 4) * 16);
    const size_t head = res_head_bytes(P.N, P.nb, MODE == 2, P.par_batched != 0);
    double *tileZ = reinterpret_cast<double *>(res_smem + head);
    double *tileU = tileZ + (size_t)P.rows_zu * 32;
    double *tileD = tileU + (size_t)P.rows_zu * 32;
    double *tileS0 = tileD + (size_t)3 * P.N * 32;
    const uint32_t z_s = (uint32_t)__cvta_generic_to_shared(tileZ);
    const uint32_t u_s = (uint32_t)__cvta_generic_to_shared(tileU);
    const uint32_t d_s = (uint32_t)__cvta_generic_to_shared(tileD);
    const int lane = threadIdx.x;
    if (lane == 0) {
        mbar_init(mbar, 1);
        mbar_init(tbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t bytes = bd_bytes + (uint32_t)(REC * P.N * 8);
        if (!P.par_batched) bytes += (uint32_t)(8 * P.nb * 8);
        mbar_expect_tx(mbar, bytes);
        bulk_g2s(fac_sbase, fac_src, (uint32_t)(REC * P.N * 8), mbar);
        if (!P.par_batched) bulk_g2s(par_sbase, P.par, (uint32_t)(8 * P.nb * 8), mbar);
        bulk_g2s((uint32_t)__cvta_generic_to_shared(bdS), P.bdesc, bd_bytes, mbar);
    }
    __syncwarp();
    mbar_wait(mbar, 0);

    FacRef<true> F;
    F.base = facS;
    F.ld = P.ld;
    F.sbase = fac_sbase;
    SmemIO io;
    io.zs = tileZ + lane; io.us = tileU + lane; io.ds = tileD + lane; io.s0s = tileS0 + lane;
    const uint32_t tile_bytes = (uint32_t)((2 * P.rows_zu + 3 * P.N) * 256);
    const int ntiles = (P.n_active + 31) >> 5;
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int col0 = tile << 5;
        const int t = col0 + lane;
        const size_t p = (size_t)t;
        const bool live = t < P.n_active && P.status[t] == ST_RUNNING;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (lane == 0) {
            // the previous tile's stores must have read their shared-memory source before it is overwritten
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_expect_tx(tbar, tile_bytes);
            for (int b = 0; b < M.zu_nbox; ++b) {
                tma_load_2d(z_s + (uint32_t)(b * M.zu_box) * 256u, &M.z, col0, b * M.zu_box, tbar);
                tma_load_2d(u_s + (uint32_t)(b * M.zu_box) * 256u, &M.u, col0, b * M.zu_box, tbar);
            }
            for (int b = 0; b < M.d_nbox; ++b)
                tma_load_2d(d_s + (uint32_t)(b * M.d_box) * 256u, &M.d, col0, b * M.d_box, tbar);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 6; ++i) tileS0[32 * i + lane] = P.s0[p + (size_t)i * P.ld];   // own column: no sync needed
        mbar_wait(tbar, tphase);
        tphase ^= 1u;

        double rho = live ? P.rho[p] : 1.0;
        double sigma = (ADAPT && live) ? P.usc[p] : 1.0;
        int it = live ? P.iters[p] : 0;
        int st = live ? ST_RUNNING : ST_MAX_ITER;
        double r_norm = 0.0, s_norm = 0.0, eps_pri = 0.0, eps_dual = 0.0;
        for (int cnt = 0; cnt < P.chunk; ++cnt) {
            const bool run = st == ST_RUNNING;
            if (!__any_sync(0xffffffffu, run)) break;
            double nr[5];
            if (MODE == 2)
                admm_iteration_dec<true, true, HAS_C, HAS_Q, ADAPT, 2, NoStaging, SmemIO>(P, p, F, bdS, par_sbase, rho, sigma, nr,
                                                                                          NoStaging(), run, io);
            else
                admm_iteration_fast<true, true, HAS_C, HAS_Q, ADAPT, SmemIO>(P, p, F, bdS, par_sbase, rho, sigma, nr, io, run);
            if (!run) continue;
            ++it;
            sigma = 1.0;
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
            if (!(isfinite(r_norm) && isfinite(s_norm))) { st = ST_NAN; continue; }
            if (r_norm < eps_pri && s_norm < eps_dual) { st = ST_CONVERGED; continue; }
            // shared factor: it does not depend on rho here (P = 0) or rho is not adapted, so no refactorisation
            if (ADAPT && (it % P.every) == 0 && it < P.max_iter && (P.until <= 0 || it <= P.until)) {
                if (r_norm > P.mu * s_norm) {
                    if (!(rho * P.tau > RHO_MAX)) { rho = rho * P.tau; sigma = P.inv_tau; }
                } else if (s_norm > P.mu * r_norm) {
                    if (!(rho * P.inv_tau < RHO_MIN)) { rho = rho * P.inv_tau; sigma = P.tau; }
                }
            }
            if (it >= P.max_iter) st = ST_MAX_ITER;
        }
        // the tile goes home: generic-proxy writes to shared memory must be visible to the async proxy first
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            for (int b = 0; b < M.zu_nbox; ++b) {
                tma_store_2d(&M.z, col0, b * M.zu_box, z_s + (uint32_t)(b * M.zu_box) * 256u);
                tma_store_2d(&M.u, col0, b * M.zu_box, u_s + (uint32_t)(b * M.zu_box) * 256u);
            }
            for (int b = 0; b < M.d_nbox; ++b)
                tma_store_2d(&M.d, col0, b * M.d_box, d_s + (uint32_t)(b * M.d_box) * 256u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (live) {
            P.iters[p] = it;
            P.rho[p] = rho;
            if (ADAPT) P.usc[p] = sigma;
            P.status[p] = st;
            P.fin[p] = r_norm;
            P.fin[p + P.ld] = s_norm;
            P.fin[p + 2 * P.ld] = eps_pri;
            P.fin[p + 3 * P.ld] = eps_dual;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace admmb
