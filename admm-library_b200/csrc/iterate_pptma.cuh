// iterate_pptma.cuh -- per-problem models (config 4: every trajectory has its own time-varying STMs): the
// stage records of the 32 problems of a warp are staged in shared memory by TMA, one stage ahead of the math.
//
// Without this the 40-46 factor entries a stage needs are 40-46 dependent-on-nothing-but-latency global loads
// per thread on the sequential stage chain: measured 212 us per iteration for a lone warp (vs 32 us with a shared
// factor) and 35 % of the HBM roofline at 16,384 problems.  Here lane 0 of every warp issues two or three
// cp.async.bulk.tensor.2d boxes ([rows of one stage] x [the warp's 32 columns]) into a two-slot ring and the
// warp reads its columns back with conflict-free LDS.64 -- "one warp per trajectory group, stage matrices staged
// in shared memory" (north_star subsystem (1), Riccati-sweep x-update).
#pragma once
#include <cuda.h>
#include "kernels.cuh"

namespace admmb {

// slot = [rows][32 lanes] doubles: 46 record rows + 6 (chat / c).  Always 52 rows: with the smaller 46-row slots two
// CTAs fit an SM, the block scheduler co-locates them while other SMs idle, and the kernel gets 8 % slower (measured).
//
// GEN = true (round 2, for the SCP loop of scp.cuh and every other per-problem model WITHOUT the in-plane / cross-track
// structure): the same ring over the generic 156-double records -- a backward slot is record rows 0..83 (K, Acl, Hinv, E)
// + chat, a forward slot K + rows 84..149 (A, B, c), 90 rows of 32 lanes -- read by admm_iteration_fast through
// staged_row() (kernels.cuh).  Without it these models ran 475 us per iteration at 4,096 problems (84-90 dependent-on-
// latency global loads per thread and stage).
__host__ __device__ constexpr int ppt_slot_bytes(bool, bool gen = false) { return (gen ? 90 : 52) * 256; }
#ifndef PPT_DEEP_PD
#define PPT_DEEP_PD 2                              // prefetch distance (stages) of the iterates in the one-warp-per-CTA form
#endif
constexpr int PPT_SLOTS = 8;                       // ring slots per CTA = warps x slots per warp: (4, 2) or (1, 8)

struct PpTmaMaps {
    // boxes of [rows] x 32 columns.  Decoupled records (fac_dec [FD*N][ld]): backward 46, forward 10 + 30, c / chat 6.
    // Generic records (fac [FS*N][ld]): backward 84, forward 18 + 66, chat 6.
    CUtensorMap mB, mF0, mF1, m6;
};

__device__ __forceinline__ void ppt_tma(uint32_t dst, const CUtensorMap *map, int col, int row, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar) : "memory");
}

// R = slots of a warp's ring (power of two): the stage records are requested R-1 stages ahead of the math.  Two slots
// (one stage ahead) are enough when many warps share an SM; a lone warp -- the narrow tail of a solve, where the slowest
// problems set the time -- needs the whole DRAM latency (~800 ns) covered by requests in flight, i.e. >= 6 stages of
// ~150 ns, so narrow working sets run one warp per CTA with an eight-slot ring (same shared memory per CTA).
template <bool HAS_C, int R, bool GEN = false>
struct PpStaging {
    const PpTmaMaps *maps;
    uint32_t slot0, bar0;                          // this warp's ring
    uint32_t lane8;
    int col0, N;
    unsigned *step;                                // uses of the ring so far (warp-uniform, lives in the kernel)
    bool lane0;

    // step s uses slot s % R with mbarrier phase (s / R) & 1
    __device__ __forceinline__ uint32_t slot(unsigned s) const { return slot0 + (s & (R - 1)) * (uint32_t)ppt_slot_bytes(HAS_C, GEN); }
    __device__ __forceinline__ uint32_t bar(unsigned s) const { return bar0 + 8u * (s & (R - 1)); }
    __device__ __forceinline__ void issue_bwd(int k, unsigned s) const
    {
        const uint32_t d = slot(s), b = bar(s);
        if (GEN) {
            mbar_expect_tx(b, (uint32_t)((F_A + (HAS_C ? 6 : 0)) * 256));
            ppt_tma(d, &maps->mB, col0, k * FS, b);
            if (HAS_C) ppt_tma(d + F_A * 256, &maps->m6, col0, k * FS + F_CHAT, b);
            return;
        }
        mbar_expect_tx(b, (uint32_t)((46 + (HAS_C ? 6 : 0)) * 256));
        ppt_tma(d, &maps->mB, col0, k * FD, b);
        if (HAS_C) ppt_tma(d + 46 * 256, &maps->m6, col0, k * FD + D_CHAT, b);
    }
    __device__ __forceinline__ void issue_fwd(int k, unsigned s) const
    {
        const uint32_t d = slot(s), b = bar(s);
        if (GEN) {
            mbar_expect_tx(b, (uint32_t)((18 + 66) * 256));
            ppt_tma(d, &maps->mF0, col0, k * FS + F_K, b);
            ppt_tma(d + 18 * 256, &maps->mF1, col0, k * FS + F_A, b);
            return;
        }
        mbar_expect_tx(b, (uint32_t)((40 + (HAS_C ? 6 : 0)) * 256));
        ppt_tma(d, &maps->mF0, col0, k * FD, b);
        ppt_tma(d + 10 * 256, &maps->mF1, col0, k * FD + D_AIN, b);
        if (HAS_C) ppt_tma(d + 40 * 256, &maps->m6, col0, k * FD + D_C, b);
    }
    // q-th stage of an iteration: backward stages N-1 .. 0, then forward stages 0 .. N-1
    __device__ __forceinline__ void issue(int q, unsigned s) const
    {
        if (q < N) issue_bwd(N - 1 - q, s);
        else if (q < 2 * N) issue_fwd(q - N, s);
    }
    template <class FR>
    __device__ __forceinline__ void iter_begin(FR &)
    {
        __syncwarp();
        if (lane0)                                 // nothing is prefetched across iterations: the factor may have
            for (int j = 0; j < R - 1; ++j) issue(j, *step + j);   // been rewritten (adaptive rho) between them
    }
    template <class FR>
    __device__ __forceinline__ void bwd_begin(int k, FR &F)
    {
        const unsigned s = *step;
        if (lane0) issue(N - 1 - k + R - 1, s + R - 1);
        mbar_wait(bar(s), (s / R) & 1u);
        F.sbase = slot(s) + lane8;
    }
    template <class FR>
    __device__ __forceinline__ void fwd_begin(int k, FR &F)
    {
        const unsigned s = *step;
        if (lane0) issue(N + k + R - 1, s + R - 1);
        mbar_wait(bar(s), (s / R) & 1u);
        F.sbase = slot(s) + lane8;
    }
    __device__ __forceinline__ void stage_end()
    {
        __syncwarp();                              // every lane has consumed the slot before it is refilled
        ++*step;
    }
};

// dynamic smem: [16 B mbarrier][par shared ? 8*nb doubles : 0][nb ints, padded][PPT_SLOTS ring mbarriers]
//               [pad to 128][PPT_SLOTS slots of ppt_slot_bytes(HAS_C)]
// W warps per CTA, each with a ring of R = PPT_SLOTS / W slots
template <bool HAS_C, bool HAS_Q, bool ADAPT, int W, bool GEN = false>
__global__ void __launch_bounds__(W * 32, 1)
k_admm_iterate_pptma(const __grid_constant__ IterParams P, const __grid_constant__ PpTmaMaps maps)
{
    extern __shared__ __align__(128) unsigned char ppt_smem[];
    unsigned char *smem_raw = ppt_smem;
    double *parS = reinterpret_cast<double *>(smem_raw + 16);
    int *bdS = reinterpret_cast<int *>(parS + (P.par_batched ? 0 : 8 * P.nb));
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t par_sbase = (uint32_t)__cvta_generic_to_shared(parS);
    const uint32_t bd_bytes = (uint32_t)(((P.nb + 3) / 4) * 16);
    const uint32_t bd_sbase = (uint32_t)__cvta_generic_to_shared(bdS);
    const uint32_t ring_bars = bd_sbase + bd_bytes;
    constexpr int R = PPT_SLOTS / W;
    const uint32_t ring = (ring_bars + PPT_SLOTS * 8 + 127u) & ~127u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        for (int i = 0; i < PPT_SLOTS; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ring_bars + 8u * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t bytes = bd_bytes;
        if (!P.par_batched) bytes += (uint32_t)(8 * P.nb * 8);
        mbar_expect_tx(mbar, bytes);
        if (!P.par_batched) bulk_g2s(par_sbase, P.par, (uint32_t)(8 * P.nb * 8), mbar);
        bulk_g2s(bd_sbase, P.bdesc, bd_bytes, mbar);
    }
    __syncthreads();
    mbar_wait(mbar, 0);

    // every lane of a warp stays in the loop (the ring protocol is warp-collective); lanes without a running
    // problem compute on a valid column of the same warp tile and never store
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int col0 = t - lane;
    if (col0 >= P.n_active) return;                                 // whole warp beyond the working set
    const bool owned = t < P.n_active;
    const size_t p = (size_t)(owned ? t : col0);
    int st = owned ? P.status[p] : ST_MAX_ITER;
    const bool was_running = st == ST_RUNNING;
    if (!__any_sync(0xffffffffu, was_running)) return;

    FacRef<false> F;
    F.base = (GEN ? P.fac : P.fac_dec) + p;
    F.ld = P.ld;
    F.sbase = 0;
    unsigned step = 0;
    PpStaging<HAS_C, R, GEN> stg;
    stg.maps = &maps;
    stg.slot0 = ring + (uint32_t)(warp * R) * ppt_slot_bytes(HAS_C, GEN);
    stg.bar0 = ring_bars + 8u * (warp * R);
    stg.lane8 = 8u * lane;
    stg.col0 = col0;
    stg.N = P.N;
    stg.step = &step;
    stg.lane0 = lane == 0;

    double rho = P.rho[p];
    double sigma = ADAPT ? P.usc[p] : 1.0;
    int it = P.iters[p];
    double r_norm = 0.0, s_norm = 0.0, eps_pri = 0.0, eps_dual = 0.0;
    // Lanes whose problem finishes in this launch keep iterating AND storing on their working column after copying
    // the final z, u, d to the home columns (kernels.cuh, k_admm_iterate: keeps the warp's stores sector-complete);
    // only when the working set is a copy (P.z_home != nullptr).
    const bool zombies = P.z_home != nullptr;
    bool zomb = false;
    auto finished = [&]() {
        if (!zombies) return;
        const size_t h = (size_t)P.orig[p];
        for (int r = 0; r < P.rows_zu; ++r) {
            P.z_home[(size_t)r * P.home_ld + h] = P.z[(size_t)r * P.ld + p];
            P.u_home[(size_t)r * P.home_ld + h] = P.u[(size_t)r * P.ld + p];
        }
        for (int r = 0; r < 3 * P.N; ++r) P.d_home[(size_t)r * P.home_ld + h] = P.d[(size_t)r * P.ld + p];
        P.snap[p] = 1;
        zomb = true;
    };
    for (int cnt = 0; cnt < P.chunk; ++cnt) {
        const bool run = st == ST_RUNNING;
        if (!__any_sync(0xffffffffu, run)) break;
        double nr[5];
        if (GEN)
            admm_iteration_fast<false, true, HAS_C, HAS_Q, ADAPT, GlobalIO, PpStaging<HAS_C, R, GEN>>(P, p, F, bdS, par_sbase, rho, zomb ? 1.0 : sigma, nr, GlobalIO(), run || zomb, stg);
        else
            admm_iteration_dec<false, true, HAS_C, HAS_Q, ADAPT, (W == 1 ? PPT_DEEP_PD : 2), PpStaging<HAS_C, R, GEN>>(P, p, F, bdS, par_sbase, rho, zomb ? 1.0 : sigma, nr, stg, run || zomb);
        if (!run) continue;
        ++it;
        sigma = 1.0;
        r_norm = sqrt(nr[0]);
        s_norm = rho * sqrt(nr[1]);
        const double nx = sqrt(nr[2]), nz = sqrt(nr[3]);
        eps_pri = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
        eps_dual = fma(P.reltol, rho * sqrt(nr[4]), P.sqrtn_abs);
        if (P.hist) {
            const size_t h = (size_t)(it - 1) * P.hist_ld + (P.orig ? (size_t)P.orig[p] : p);
            P.hist[h] = r_norm;
            P.hist[h + P.hist_stride] = s_norm;
            P.hist[h + 2 * P.hist_stride] = eps_pri;
            P.hist[h + 3 * P.hist_stride] = eps_dual;
            P.hist[h + 4 * P.hist_stride] = rho;
        }
        if (!(isfinite(r_norm) && isfinite(s_norm))) { st = ST_NAN; finished(); continue; }
        if (r_norm < eps_pri && s_norm < eps_dual) { st = ST_CONVERGED; finished(); continue; }
        if (ADAPT && (it % P.every) == 0 && it < P.max_iter && (P.until <= 0 || it <= P.until)) {
            bool ch = false;
            if (r_norm > P.mu * s_norm) {
                if (!(rho * P.tau > RHO_MAX)) { rho = rho * P.tau; sigma = P.inv_tau; ch = true; }
            } else if (s_norm > P.mu * r_norm) {
                if (!(rho * P.inv_tau < RHO_MIN)) { rho = rho * P.inv_tau; sigma = P.tau; ch = true; }
            }
            if (ch && P.has_P) {
                const size_t off = P.raw_batched ? p : 0, ldr = P.raw_batched ? P.ld : 1;
                int bad = riccati_factor_dev(P.N, P.rawA + off, P.rawB + off, P.rawc ? P.rawc + off : nullptr,
                                             P.rawQ ? P.rawQ + off : nullptr, P.rawR ? P.rawR + off : nullptr,
                                             ldr, rho, bdS, P.fac_rw + p, P.ld);
                if (owned && t < P.n_real) atomicAdd(P.refac_count, 1ULL);
                if (bad) { st = ST_NAN; finished(); continue; }
                if (!GEN) pack_decoupled_dev(P.N, P.fac_rw + p, P.ld, P.fac_dec_rw + p, P.ld);
                __threadfence();                                   // the next iteration's TMA reads must see the new record
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
        }
        if (it >= P.max_iter) { st = ST_MAX_ITER; finished(); }
    }
    if (was_running) {
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

}  // namespace admmb
