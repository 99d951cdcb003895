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
//   warp 5  NORMS               trails the prox warps block by block (polling their progress counters): the five norm
//                               accumulators, each summed over the blocks in the oracle's order.
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
#ifdef WG_TIMING
#include <cstdio>
#endif
#include "kernels.cuh"

namespace admmb {

constexpr int WG_WARPS = 8;
// PP (per-problem models, config 4): the stage records of the tile's problems are streamed through a ring of WG_PP_R
// shared-memory slots by TMA (one box of [record rows] x [TW columns] per sweep stage, issued by warp 4, which is
// otherwise idle) instead of being read from a shared factor table; full / empty mbarriers per slot.
#ifndef WG_PP_RING
#define WG_PP_RING 8
#endif
constexpr int WG_PP_R = WG_PP_RING;               // ring slots (power of two; -DWG_PP_RING=n: developer builds)
constexpr int WG_PP_FULL_COUNT = 1;               // full barrier: the producer's expect_tx arrival (+ the bulk copy's bytes)
constexpr int WG_PP_ROWS = D_AIN;                // rows of a slot: a backward stage needs record rows 0..45, a forward stage 10 + 30
struct WgPpMaps {
    const double *blk;                           // tile-blocked copy of the records (k_wgpp_block)
};
// Tile-blocked copy of the per-problem stage records: [tile][stage][WG_PP_BLK_ROWS rows][TW columns], rows 0..45 = what a
// backward stage reads (record rows 0..45: K, Acl, Hinv, E), rows 46..85 = what a forward stage reads (K again: 10 rows, then
// A, B: record rows 46..75).  One ring slot is then ONE contiguous bulk copy (cp.async.bulk, 8.8 KB at TW = 24) instead of a
// TMA box of 46 strided rows.  Rows are interleaved in PAIRS -- element (r, c) at ((r / 2) * TW + c) * 2 + (r & 1) -- so that
// a lane reads two consecutive record entries of its problem with one LDS.128 (every operand group starts on an even row).
constexpr int WG_PP_BLK_ROWS = D_AIN + 40;
static __global__ void k_wgpp_block(const double *__restrict__ fac_dec, size_t ld, int N, int n_active, int TW, double *__restrict__ blk)
{
    const int tile = blockIdx.x, k = blockIdx.y;
    const double *src = fac_dec + (size_t)(k * FD) * ld;
    double *dst = blk + ((size_t)tile * N + k) * WG_PP_BLK_ROWS * TW;
    for (int i = threadIdx.x; i < WG_PP_BLK_ROWS * TW; i += blockDim.x) {
        const int r = i / TW, c = i - r * TW;
        const int rr = r < D_AIN ? r : (r < D_AIN + 10 ? r - D_AIN : r - 10);     // record row
        const int col = min(tile * TW + c, n_active - 1);                        // columns past the batch: copies of its last problem
        dst[((size_t)(r >> 1) * TW + c) * 2 + (r & 1)] = __ldcg(src + (size_t)rr * ld + col);
    }
}
constexpr int WG_MIN_TW = 8;                     // narrower tiles waste more than 3/4 of every warp: not worth it
constexpr int WG_PROX = 4;                       // prox warps

struct WgLayout {                                // byte offsets inside the dynamic shared memory of one CTA
    int TW;                                      // tile width: problems per CTA
    int rz, rd, rg;                              // doubles per problem in the z / u, d and g arrays (odd: see below)
    size_t bars, prog, fac, par, typ, z, u, d, g, s0, nrm, rbar, ring, total;
};

// Tile arrays are stored PROBLEM-major: the rows of one problem are contiguous, lane p starts at p * pitch.  A row is
// then a compile-time offset from a per-lane pointer (no index arithmetic on the sequential chains), and an odd pitch
// keeps a warp's 64-bit accesses conflict-free (16 lanes of a half-warp hit 16 distinct bank pairs).
// fw = doubles per stage of the factor copy in shared memory: the whole packed record (FD), or -- time-invariant dynamics,
// whose A_k, B_k live in the chain warps' registers -- only K, Acl, Hinv, E (the first D_AIN = 46 doubles): 18 KB instead of
// 35 KB at N = 50, 37 KB instead of 70 KB at N = 100, i.e. tiles of 32 instead of 30 and of 15 instead of 12 problems.
__host__ __device__ inline WgLayout wg_layout(int N, int rows_zu, int TW, int fw = FD, bool pp = false)
{
    WgLayout L;
    L.TW = TW;
    L.rz = rows_zu | 1;
    L.rd = (3 * N) | 1;
    L.rg = (6 * N + 12) | 1;                     // per stage a_k (3) + ds_k (3); then s_N (6) and the terminal ds (6)
    const int nsb = rows_zu / 3;                 // split blocks: N controls + the split terminal blocks
    size_t o = 16;                               // [0, 8): mbarrier of the factor copy
    L.bars = o; o += 8 * (size_t)(N + 1);        // bC2[0..N]: a_k / s_N published by both chain warps
    o = (o + 15) / 16 * 16;
    L.prog = o; o += 16;                         // blocks finished by each prox warp since the launch began (4 x u32)
    L.fac = o; o += sizeof(double) * (size_t)fw * (size_t)N;
    L.par = o; o += sizeof(double) * 8 * (size_t)nsb;          // parameter table, compact: one record per split block
    L.typ = o; o += sizeof(int) * (size_t)((nsb + 3) / 4 * 4);
    const size_t col = sizeof(double) * (size_t)TW;
    L.z = o; o += col * L.rz;
    L.u = o; o += col * L.rz;
    L.d = o; o += col * L.rd;
    L.g = o; o += col * L.rg;
    L.s0 = o; o += col * 7;
    L.nrm = o; o += col * 15;                    // two sets of five sums, alternating by iteration parity; five roots
    L.rbar = L.ring = 0;
    if (pp) {                                    // full[R], empty[R] mbarriers, then R slots of WG_PP_ROWS x TW doubles (128-byte aligned)
        o = (o + 15) / 16 * 16;
        L.rbar = o; o += 16 * (size_t)WG_PP_R;
        o = (o + 127) / 128 * 128;
        L.ring = o; o += (size_t)WG_PP_R * WG_PP_ROWS * col;
    }
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

// Factor entries `off .. off+3` of the current stage.  Shared table: fk = address of the stage's record, one broadcast
// 128-bit load per pair.  PP: fk = address of this lane's entry of row pair 0 in the ring slot, one 128-bit load per pair of
// entries (rows are interleaved in pairs, see k_wgpp_block); slot row of record offset `off`: off below 46, off - 36 above
// (a forward slot is K (10 rows) followed by A, B (30 rows)); rp = bytes per row.
template <bool PP>
__device__ __forceinline__ void wgf_ld2(uint32_t fk, int off, uint32_t rp, double &x, double &y)
{
    static_assert((D_KIN | D_KC | D_ACLIN | D_ACLC | D_HIN | D_HC | D_EIN | D_EC | D_AIN | D_AC | D_BIN | D_BC) % 2 == 0, "row pairs");
    if (PP) wg_ld2(fk + (uint32_t)((off < D_AIN ? off : off - 36) >> 1) * (2u * rp), x, y);      // fk: this lane's 16 bytes of row pair 0
    else wg_ld2(fk + (uint32_t)off * 8u, x, y);
}
template <bool PP>
__device__ __forceinline__ void wgf_ld4(uint32_t fk, int off, uint32_t rp, double (&r)[4])
{
    wgf_ld2<PP>(fk, off, rp, r[0], r[1]);
    wgf_ld2<PP>(fk, off + 2, rp, r[2], r[3]);
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

// One non-blocking test of a barrier phase (the result is used a stage later: see pre_ok in the kernel)
__device__ __forceinline__ uint32_t wg_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

// Same hand-over, polled with test_wait: a try_wait that has to block suspends the warp and wakes it up late; the prox warps
// sit right behind the chain, where that wake-up latency goes straight into the length of an iteration.
__device__ __forceinline__ void wg_wait_spin(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#ifdef WG_SPIN_BACKOFF
        if (!ok) __nanosleep(WG_SPIN_BACKOFF);
#endif
        if (!ok && ++spins > (1 << 24)) __trap();
    } while (!ok);
}

// Stage loop with the operands of the next stage loaded (into registers) BEFORE the current stage's arithmetic and
// stores: a warp issues in order, so without this every stage would start by waiting out its own load latency.
// Stage j = 0 .. count-1 is k = k0 + j * stride.  Branch-free body: the prefetch past the end re-reads the last stage.
// EXACT (PP): every stage is loaded exactly once (a load consumes a ring slot), straight before its arithmetic.
template <class In, bool EXACT = false, class LoadFn, class WorkFn>
__device__ __forceinline__ void wg_stage_loop(const int k0, const int stride, const int count, LoadFn load, WorkFn work)
{
    if (count <= 0) return;
    if (EXACT) {
        // PP: the operands already wait in shared memory (the TMA ring IS the prefetch), so one register copy of a stage is
        // enough.  Two copies (80 doubles) next to the chain state took all 255 registers: spills and 44 register moves
        // inside every pair of stages, 240 instead of 105 instructions per stage on a warp that issues one per 2.3 cycles.
        int k = k0;
        for (int j = 0; j < count; ++j, k += stride) {
            In a;
            load(k, a);
            work(k, a);
        }
        return;
    }
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

#ifdef WG_TIMING
// developer build: absolute SM-clock stamps of one iteration (the 8th of the launch) of CTA 0, printed when the launch ends
__device__ long long wg_tlog[WG_WARPS][8];
__device__ long long wg_klog[3][64];               // per stage: a_k published (in-plane chain), block k done (prox), block k accumulated
#define WG_T(idx) do { if (blockIdx.x == 0 && (tid & 31) == 0 && cnt == 7) wg_tlog[warp][idx] = clock64(); } while (0)
#define WG_TK(row, k) do { if (blockIdx.x == 0 && (tid & 31) == 0 && cnt == 7 && (k) < 64) wg_klog[row][k] = clock64(); } while (0)
#else
#define WG_T(idx) do { } while (0)
#define WG_TK(row, k) do { } while (0)
#endif

// TI: A_k, B_k do not depend on the stage (time-invariant dynamics, e.g. Clohessy-Wiltshire with a fixed step): the chain
// warps keep them in registers for the whole launch instead of re-reading 30 doubles per stage on the forward sweep
// PTW (PP only): the tile width as a compile-time constant.  A ring slot is [record rows] x [TW columns], so the row pitch
// of every operand load of the chain warps is TW * 8 bytes: with a run-time TW each of the 46 + 40 loads of a stage pair
// needed its own IMAD and address register (255 registers, spills inside the sweep loops, 240 instructions per stage);
// as a constant they are immediates off one base register.
template <bool ADAPT, bool TI, bool PP = false, int PTW = 0>
__global__ void __launch_bounds__(WG_WARPS * 32, 1)
k_admm_iterate_wg(const __grid_constant__ IterParams P, const int TW_arg, const __grid_constant__ WgPpMaps maps)
{
    static_assert(!(TI && PP), "per-problem models are not time-invariant");
    static_assert(PP == (PTW > 0) && PTW % 8 == 0, "the streamed-record form is instantiated per tile width (8, 16, 24, 32)");
    const int TW = PTW > 0 ? PTW : TW_arg;
    extern __shared__ __align__(128) unsigned char wg_smem[];
    const int N = P.N;
    constexpr int FW = PP ? 0 : (TI ? D_AIN : FD);       // stage pitch of the factor copy in shared memory (see wg_layout)
    const WgLayout L = wg_layout(N, P.rows_zu, TW, FW, PP);
    const int tid = threadIdx.x, warp = tid >> 5;
    // lanes beyond the tile width are exact clones of the tile's last lane (same problem, same state, same values
    // written to the same addresses), so no access has to be predicated on the lane
    const int lane = min(tid & 31, TW - 1);
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(wg_smem);
    const uint32_t mbar = sm0;
    const uint32_t bC2 = sm0 + (uint32_t)L.bars, prog_s = sm0 + (uint32_t)L.prog;
    const uint32_t fac_s = sm0 + (uint32_t)L.fac, par_s = sm0 + (uint32_t)L.par, typ_s = sm0 + (uint32_t)L.typ;
    // this lane's problem in every tile array (byte addresses; row r of an array is at base + 8 r)
    const uint32_t zs = sm0 + (uint32_t)L.z + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t us = sm0 + (uint32_t)L.u + (uint32_t)(lane * L.rz) * 8u;
    const uint32_t ds = sm0 + (uint32_t)L.d + (uint32_t)(lane * L.rd) * 8u;
    const uint32_t gs = sm0 + (uint32_t)L.g + (uint32_t)(lane * L.rg) * 8u;
    const uint32_t s0s = sm0 + (uint32_t)L.s0 + (uint32_t)(lane * 7) * 8u;
    const uint32_t nrs0 = sm0 + (uint32_t)L.nrm + (uint32_t)(lane * 15) * 8u;
    const uint32_t gN = gs + (uint32_t)(6 * N) * 8u;      // s_N in natural order (6 rows), then the terminal blocks' ds (6 rows)

    // roles
    const bool chain_in = warp == 0, chain_c = warp == 1, norms = warp == 5;
    const int prox = warp == 2 ? 0 : warp == 3 ? 1 : warp == 6 ? 2 : warp == 7 ? 3 : -1;

    // ---- once per CTA: factor (one TMA bulk copy), hand-over barriers, compact parameter table and block types
    // PP ring: full[s] at rbar + 8 s (one arrival + the TMA bytes), empty[s] at rbar + 8 (R + s) (both chain warps)
    const uint32_t rfull = sm0 + (uint32_t)L.rbar, rempty = rfull + 8u * WG_PP_R, ring_s = sm0 + (uint32_t)L.ring;
    const uint32_t rp = (uint32_t)TW * 8u;               // row pitch of a slot
    const uint32_t slot_bytes = (uint32_t)WG_PP_ROWS * rp;
    // PP, chain warps: the NEXT slot's full barrier is tested right after a stage's operands have been read, so the ~100-cycle
    // round trip of the test runs under that stage's arithmetic instead of in front of the next stage's loads
    uint32_t pre_ok = 0;
    uint32_t cstep = 0, pstep = 0;                       // ring steps consumed (chain warps) / produced (warp 4) since the launch began
    if (tid == 0) {
        mbar_init(mbar, 1);
        for (int k = 0; k <= N; ++k) mbar_init(bC2 + 8u * k, 2);      // both chain warps
        for (int w = 0; w < WG_PROX; ++w) asm volatile("st.shared.u32 [%0], %1;" ::"r"(prog_s + 4u * w), "r"(0) : "memory");
        if (PP) {
            for (int i = 0; i < WG_PP_R; ++i) { mbar_init(rfull + 8u * i, WG_PP_FULL_COUNT); mbar_init(rempty + 8u * i, 2); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        mbar_expect_tx(mbar, (uint32_t)(FW * N * 8));
        if (PP) {
            // nothing to copy: the arrival of expect_tx(0) completes the phase
        } else if (TI) {   // K, Acl, Hinv, E of every stage: N copies of 368 bytes (both ends 16-byte aligned: 704 k and 368 k)
            for (int k = 0; k < N; ++k)
                bulk_g2s(fac_s + (uint32_t)(k * FW) * 8u, P.fac_dec + (size_t)k * FD, (uint32_t)(FW * 8), mbar);
        } else {
            bulk_g2s(fac_s, P.fac_dec, (uint32_t)(FD * N * 8), mbar);
        }
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

    double Ai[4][4], Bi[4][2], Ac[2][2], Bc[2];          // TI: the stage-invariant dynamics of this warp's chain
    if (TI && chain_in) {                                // from the global record of stage 0 (not part of the compact copy)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < 4; ++c) Ai[r][c] = __ldg(P.fac_dec + D_AIN + 4 * r + c);
            Bi[r][0] = __ldg(P.fac_dec + D_BIN + 2 * r);
            Bi[r][1] = __ldg(P.fac_dec + D_BIN + 2 * r + 1);
        }
    }
    if (TI && chain_c) {
        Ac[0][0] = __ldg(P.fac_dec + D_AC); Ac[0][1] = __ldg(P.fac_dec + D_AC + 1);
        Ac[1][0] = __ldg(P.fac_dec + D_AC + 2); Ac[1][1] = __ldg(P.fac_dec + D_AC + 3);
        Bc[0] = __ldg(P.fac_dec + D_BC); Bc[1] = __ldg(P.fac_dec + D_BC + 1);
    }

    const size_t ld = P.ld;
    const int ntiles = (P.n_active + TW - 1) / TW;
    uint32_t ph = 0;                                     // parity of the hand-over barriers: flips once per iteration
    uint32_t pcount = 0;                                 // prox warps: blocks finished since the launch began
    uint32_t nbase[WG_PROX] = {0, 0, 0, 0};              // norm warp: the prox warps' counts when the current iteration began
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
            WG_T(0);

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
                    wg_stage_loop<In, PP>(N - 1, -1, N,
                        [&](int k, In &in) {
                            uint32_t fk, sl = 0;
                            if (PP) {                                   // this stage's records: wait for the slot
                                const uint32_t o = (uint32_t)(3 * k) * 8u;  // (z, u first: they do not wait for the ring)
#pragma unroll
                                for (int e = 0; e < 2; ++e) { in.z[e] = wg_ld(zs + o + 8u * e); in.u[e] = wg_ld(us + o + 8u * e); }
                                sl = cstep & (WG_PP_R - 1);
                                if (!pre_ok) wg_wait_spin(rfull + 8u * sl, (cstep / WG_PP_R) & 1u);
                                fk = ring_s + sl * slot_bytes + (uint32_t)lane * 16u;
                            } else {
                                fk = fac_s + (uint32_t)(k * FW) * 8u;
                            }
                            wgf_ld4<PP>(fk, D_KIN, rp, in.k0);
                            wgf_ld4<PP>(fk, D_KIN + 4, rp, in.k1);
#pragma unroll
                            for (int l = 0; l < 4; ++l) wgf_ld4<PP>(fk, D_ACLIN + 4 * l, rp, in.ac[l]);
                            wgf_ld2<PP>(fk, D_HIN, rp, in.h0[0], in.h0[1]);
                            wgf_ld2<PP>(fk, D_HIN + 2, rp, in.h1[0], in.h1[1]);
                            wgf_ld4<PP>(fk, D_EIN, rp, in.e0);
                            wgf_ld4<PP>(fk, D_EIN + 4, rp, in.e1);
                            if (PP) {                                   // operands are in registers: hand the slot back
                                wg_arrive(rempty + 8u * sl);
                                ++cstep;
                                pre_ok = wg_test(rfull + 8u * (cstep & (WG_PP_R - 1)), (cstep / WG_PP_R) & 1u);
                            }
                            if (!PP) {
                                const uint32_t o = (uint32_t)(3 * k) * 8u;
#pragma unroll
                                for (int e = 0; e < 2; ++e) { in.z[e] = wg_ld(zs + o + 8u * e); in.u[e] = wg_ld(us + o + 8u * e); }
                            }
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
                WG_T(1);
                // ---- forward sweep: a_k = d_k + K_k s_k (published for the prox warps), s_{k+1} = A_k s_k + B_k a_k
                double si[4] = {wg_ld(s0s), wg_ld(s0s + 8), wg_ld(s0s + 24), wg_ld(s0s + 32)};
                {
                    struct In { double d[2], kn[2][4], an[4][4], bn[4][2]; };
                    wg_stage_loop<In, PP>(0, 1, N,
                        [&](int k, In &o) {
                            const uint32_t dk = ds + (uint32_t)(3 * k) * 8u;
                            uint32_t fk, sl = 0;
                            if (PP) {                                   // this stage's records: wait for the slot
                                o.d[0] = wg_ld(dk);
                                o.d[1] = wg_ld(dk + 8);
                                sl = cstep & (WG_PP_R - 1);
                                if (!pre_ok) wg_wait_spin(rfull + 8u * sl, (cstep / WG_PP_R) & 1u);
                                fk = ring_s + sl * slot_bytes + (uint32_t)lane * 16u;
                            } else {
                                fk = fac_s + (uint32_t)(k * FW) * 8u;
                            }
                            wgf_ld4<PP>(fk, D_KIN, rp, o.kn[0]);
                            wgf_ld4<PP>(fk, D_KIN + 4, rp, o.kn[1]);
                            if (!TI) {
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    wgf_ld4<PP>(fk, D_AIN + 4 * r, rp, o.an[r]);
                                    wgf_ld2<PP>(fk, D_BIN + 2 * r, rp, o.bn[r][0], o.bn[r][1]);
                                }
                            }
                            if (PP) {                                   // operands are in registers: hand the slot back
                                wg_arrive(rempty + 8u * sl);
                                ++cstep;
                                pre_ok = wg_test(rfull + 8u * (cstep & (WG_PP_R - 1)), (cstep / WG_PP_R) & 1u);
                            }
                            if (!PP) {
                                o.d[0] = wg_ld(dk);
                                o.d[1] = wg_ld(dk + 8);
                            }
                        },
                        [&](int k, const In &o) {
                            double a0 = o.d[0], a1 = o.d[1];
#pragma unroll
                            for (int i = 0; i < 4; ++i) { a0 = fma(o.kn[0][i], si[i], a0); a1 = fma(o.kn[1][i], si[i], a1); }
                            const uint32_t gk = gs + (uint32_t)(6 * k) * 8u;
                            wg_st(gk, a0);
                            wg_st(gk + 8, a1);
                            wg_arrive(bC2 + 8u * k);
                            WG_TK(0, k);
                            const auto &an = TI ? Ai : o.an;
                            const auto &bn = TI ? Bi : o.bn;
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
#pragma unroll
                            for (int r = 0; r < 4; ++r) si[r] = ni[r];
                        });
                }
                wg_st(gN, si[0]); wg_st(gN + 8, si[1]); wg_st(gN + 24, si[2]); wg_st(gN + 32, si[3]);   // s_N, natural order
                wg_arrive(bC2 + 8u * N);
                WG_T(2);
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
                    wg_stage_loop<In, PP>(N - 1, -1, N,
                        [&](int k, In &in) {
                            uint32_t fk, sl = 0;
                            if (PP) {                                   // this stage's records: wait for the slot
                                const uint32_t o = (uint32_t)(3 * k + 2) * 8u;
                                in.z = wg_ld(zs + o);
                                in.u = wg_ld(us + o);
                                sl = cstep & (WG_PP_R - 1);
                                if (!pre_ok) wg_wait_spin(rfull + 8u * sl, (cstep / WG_PP_R) & 1u);
                                fk = ring_s + sl * slot_bytes + (uint32_t)lane * 16u;
                            } else {
                                fk = fac_s + (uint32_t)(k * FW) * 8u;
                            }
                            wgf_ld2<PP>(fk, D_KC, rp, in.kc[0], in.kc[1]);
                            wgf_ld2<PP>(fk, D_ACLC, rp, in.a0[0], in.a0[1]);
                            wgf_ld2<PP>(fk, D_ACLC + 2, rp, in.a1[0], in.a1[1]);
                            wgf_ld2<PP>(fk, D_HC, rp, in.hc[0], in.hc[1]);
                            wgf_ld2<PP>(fk, D_EC, rp, in.ec[0], in.ec[1]);
                            if (PP) {                                   // operands are in registers: hand the slot back
                                wg_arrive(rempty + 8u * sl);
                                ++cstep;
                                pre_ok = wg_test(rfull + 8u * (cstep & (WG_PP_R - 1)), (cstep / WG_PP_R) & 1u);
                            }
                            if (!PP) {
                                const uint32_t o = (uint32_t)(3 * k + 2) * 8u;
                                in.z = wg_ld(zs + o);
                                in.u = wg_ld(us + o);
                            }
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
                WG_T(1);
                double sc0 = wg_ld(s0s + 16), sc1 = wg_ld(s0s + 40);
                {
                    struct In { double kc[2], a0[2], a1[2], bc[2], d2; };
                    wg_stage_loop<In, PP>(0, 1, N,
                        [&](int k, In &in) {
                            uint32_t fk, sl = 0;
                            if (PP) {                                   // this stage's records: wait for the slot
                                in.d2 = wg_ld(ds + (uint32_t)(3 * k + 2) * 8u);
                                sl = cstep & (WG_PP_R - 1);
                                if (!pre_ok) wg_wait_spin(rfull + 8u * sl, (cstep / WG_PP_R) & 1u);
                                fk = ring_s + sl * slot_bytes + (uint32_t)lane * 16u;
                            } else {
                                fk = fac_s + (uint32_t)(k * FW) * 8u;
                            }
                            wgf_ld2<PP>(fk, D_KC, rp, in.kc[0], in.kc[1]);
                            if (!TI) {
                                wgf_ld2<PP>(fk, D_AC, rp, in.a0[0], in.a0[1]);
                                wgf_ld2<PP>(fk, D_AC + 2, rp, in.a1[0], in.a1[1]);
                                wgf_ld2<PP>(fk, D_BC, rp, in.bc[0], in.bc[1]);
                            }
                            if (PP) {                                   // operands are in registers: hand the slot back
                                wg_arrive(rempty + 8u * sl);
                                ++cstep;
                                pre_ok = wg_test(rfull + 8u * (cstep & (WG_PP_R - 1)), (cstep / WG_PP_R) & 1u);
                            }
                            if (!PP) in.d2 = wg_ld(ds + (uint32_t)(3 * k + 2) * 8u);
                        },
                        [&](int k, const In &in) {
                            double a2 = fma(in.kc[0], sc0, in.d2);
                            a2 = fma(in.kc[1], sc1, a2);
                            wg_st(gs + (uint32_t)(6 * k + 2) * 8u, a2);
                            wg_arrive(bC2 + 8u * k);
                            const auto &a0r = TI ? Ac[0] : in.a0;
                            const auto &a1r = TI ? Ac[1] : in.a1;
                            const auto &bcr = TI ? Bc : in.bc;
                            double n0 = a0r[0] * sc0;
                            n0 = fma(a0r[1], sc1, n0);
                            n0 = fma(bcr[0], a2, n0);
                            double n1 = a1r[0] * sc0;
                            n1 = fma(a1r[1], sc1, n1);
                            n1 = fma(bcr[1], a2, n1);
                            sc0 = n0; sc1 = n1;
                        });
                }
                wg_st(gN + 16, sc0); wg_st(gN + 40, sc1);
                wg_arrive(bC2 + 8u * N);
                WG_T(2);
            } else if (prox >= 0) {
                // ================= prox warps: relaxation, prox, dual ascent of block k behind the forward sweep
                struct In { double x[3], z[3], u[3], pr[8]; int type; };
                // block `slot` (compact index): everything that does not depend on the chain (parameters, type, z, u) is
                // loaded BEFORE the wait on the chain's barrier, x = a_k (at xa) after it; ds = z+ - z goes to da
                auto load_own = [&](int slot, In &in) {
                    const uint32_t pa = par_s + (uint32_t)slot * 64u, o = (uint32_t)(3 * slot) * 8u;
                    wg_ld2(pa, in.pr[0], in.pr[1]); wg_ld2(pa + 16, in.pr[2], in.pr[3]);
                    wg_ld2(pa + 32, in.pr[4], in.pr[5]); wg_ld2(pa + 48, in.pr[6], in.pr[7]);
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(in.type) : "r"(typ_s + 4u * slot) : "memory");
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        in.z[e] = wg_ld(zs + o + 8u * e);
                        in.u[e] = wg_ld(us + o + 8u * e);
                    }
                };
                auto update_block = [&](int slot, uint32_t xa, uint32_t da, In &in) {
#pragma unroll
                    for (int e = 0; e < 3; ++e) in.x[e] = wg_ld(xa + 8u * e);
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
                // Blocks 0 .. N+1 go round the prox warps: control block k (compact index k in the fast pattern; x = a_k at g
                // rows 6k..6k+2, ds to 6k+3..), then the two terminal blocks (x = s_N).  After each block the warp publishes
                // its running count (release), which is what the norm warp polls.
                for (int k = prox; k < N + 2; k += WG_PROX) {
                    In in;
                    if (k < N) {
                        load_own(k, in);
                        wg_wait_spin(bC2 + 8u * k, ph);
                        update_block(k, gs + (uint32_t)(6 * k) * 8u, gs + (uint32_t)(6 * k + 3) * 8u, in);
                    } else {
                        const int tb = k - N;
                        const int de = tb == 0 ? de_t0 : de_t1;
                        if ((de & 0xff) != BLK_NONE) load_own(de >> 8, in);
                        wg_wait_spin(bC2 + 8u * N, ph);
                        if ((de & 0xff) != BLK_NONE)
                            update_block(de >> 8, gN + (uint32_t)(3 * tb) * 8u, gN + (uint32_t)(6 + 3 * tb) * 8u, in);
                    }
                    ++pcount;
                    __syncwarp();
                    if ((tid & 31) == 0)
                        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(prog_s + 4u * prox), "r"(pcount) : "memory");
                    WG_TK(1, k);
                }
                WG_T(3);
            } else if (PP && warp == 4) {
                // ================= per-problem models: the records of the 2 N sweep stages of this iteration, in the order
                // the chain warps consume them (backward N-1 .. 0, forward 0 .. N-1), WG_PP_R - 1 stages ahead of them at most
                for (int q = 0; q < 2 * N; ++q) {
                    const uint32_t sl = pstep & (WG_PP_R - 1);
                    wg_wait(rempty + 8u * sl, ((pstep / WG_PP_R) & 1u) ^ 1u);      // both chain warps are done with the slot
                    if ((tid & 31) == 0) {
                        const uint32_t dst = ring_s + sl * slot_bytes, bar = rfull + 8u * sl;
                        const bool bwd = q < N;
                        const int k = bwd ? N - 1 - q : q - N;
                        const double *src = maps.blk + (((size_t)tile * N + k) * WG_PP_BLK_ROWS + (bwd ? 0 : D_AIN)) * TW;
                        const uint32_t bytes = (bwd ? (uint32_t)WG_PP_ROWS : 40u) * rp;
                        mbar_expect_tx(bar, bytes);
                        bulk_g2s(dst, src, bytes, bar);
                    }
                    ++pstep;
                }
            } else if (norms) {
                // ================= the five norm accumulators, blocks in the oracle's order, behind the prox warps
                double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
                struct In { double x[3], z[3], u[3], dd[3]; };
                auto load_block = [&](uint32_t xa, uint32_t za, uint32_t ua, uint32_t da, In &in) {
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        in.x[e] = wg_ld(xa + 8u * e); in.z[e] = wg_ld(za + 8u * e);
                        in.u[e] = wg_ld(ua + 8u * e); in.dd[e] = wg_ld(da + 8u * e);
                    }
                };
                auto add_block = [&](const In &in) {
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double dr = in.x[e] - in.z[e];
                        rr = fma(dr, dr, rr);
                        ss = fma(in.dd[e], in.dd[e], ss);
                        xx = fma(in.x[e], in.x[e], xx);
                        zz = fma(in.z[e], in.z[e], zz);
                        uu = fma(in.u[e], in.u[e], uu);
                    }
                };
                // `ready` = number of leading blocks known to be finished, from the prox warps' counters: warp w has done
                // its first c_w blocks of this iteration, i.e. every block below 4 c_w + w of its residue class
                int ready = 0;
                auto need = [&](int k) {
                    int spins = 0;
                    while (ready <= k) {
                        uint32_t c[WG_PROX];
#pragma unroll
                        for (int w = 0; w < WG_PROX; ++w)
                            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(c[w]) : "r"(prog_s + 4u * w) : "memory");
                        int r = 1 << 30;
#pragma unroll
                        for (int w = 0; w < WG_PROX; ++w) r = min(r, 4 * (int)(c[w] - nbase[w]) + w);
                        // warp-uniform run boundaries: loads of a counter that is bumped meanwhile can return different
                        // values to different lanes (a wide load is served a quarter-warp at a time), and lanes beyond the
                        // tile width are copies of its last lane that must stay in lockstep with it
                        ready = __shfl_sync(0xffffffffu, r, 0);
                        if (++spins > (1 << 22)) __trap();
                    }
                };
                // Runs of finished blocks: one poll tells how far the prox warps have got, then blocks k .. kend-1 go through
                // a branch-free, software-pipelined loop (block k + 1 is loaded before block k is accumulated).  A lone warp
                // pays 10-15 cycles for every taken branch, so nothing is tested per block.
                {
                    auto load_k = [&](int k, In &in) {
                        load_block(gs + (uint32_t)(6 * k) * 8u, zs + (uint32_t)(3 * k) * 8u, us + (uint32_t)(3 * k) * 8u,
                                   gs + (uint32_t)(6 * k + 3) * 8u, in);
                    };
                    int k = 0;
                    while (k < N) {
                        need(k);
                        const int kend = min(ready, N);
                        In a, b;
                        load_k(k, a);
                        for (; k + 2 < kend; k += 2) {
                            load_k(k + 1, b);
                            add_block(a);
                            load_k(k + 2, a);
                            add_block(b);
                        }
                        if (k + 1 < kend) { load_k(k + 1, b); add_block(a); add_block(b); k += 2; }
                        else { add_block(a); k += 1; }
                        WG_TK(2, k - 1);
                    }
                }
                WG_T(4);
                need(N + 1);
#pragma unroll
                for (int tb = 0; tb < 2; ++tb) {
                    const int de = tb == 0 ? de_t0 : de_t1;
                    if ((de & 0xff) == BLK_NONE) continue;
                    const uint32_t o = (uint32_t)(3 * (de >> 8)) * 8u;
                    In in;
                    load_block(gN + (uint32_t)(3 * tb) * 8u, zs + o, us + o, gN + (uint32_t)(6 + 3 * tb) * 8u, in);
                    add_block(in);
                }
#pragma unroll
                for (int w = 0; w < WG_PROX; ++w) nbase[w] += (uint32_t)((N + 2 - w + WG_PROX - 1) / WG_PROX);
                wg_st(nrs, rr); wg_st(nrs + 8, ss); wg_st(nrs + 16, xx); wg_st(nrs + 24, zz); wg_st(nrs + 32, uu);
            }
            WG_T(5);
            ph ^= 1u;
            __syncthreads();
            WG_T(6);

            // ================= stopping test and rho update: every warp, for its own copy of the lane's state.  The five
            // square roots are the long part (dependent Newton chains, ~150 cycles each): warps 0-4 take one each
            if (warp < 5) wg_st(nrs0 + 80u + 8u * warp, sqrt(wg_ld(nrs + 8u * warp)));
            __syncthreads();
            if (run) {
                ++it;
                sigma = 1.0;
                const uint32_t rts = nrs0 + 80u;
                r_norm = wg_ld(rts);
                s_norm = rho * wg_ld(rts + 8);
                const double nx = wg_ld(rts + 16), nz = wg_ld(rts + 24);
                eps_pri = fma(P.reltol, nx > nz ? nx : nz, P.sqrtn_abs);
                eps_dual = fma(P.reltol, rho * wg_ld(rts + 32), P.sqrtn_abs);
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
            WG_T(7);
        }

#ifdef WG_TIMING
        if (blockIdx.x == 0 && tid == 0) {
            const char *names[8] = {"start", "bwd_done", "fwd_done", "prox_done", "norm_ctl", "role_done", "barrier", "stop_done"};
            for (int w = 0; w < WG_WARPS; ++w)
                for (int i = 0; i < 8; ++i)
                    if (wg_tlog[w][i]) printf("wgt warp %d %-10s %lld\n", w, names[i], wg_tlog[w][i] - wg_tlog[0][0]);
            for (int k = 0; k < 64 && k < N + 2; ++k)
                printf("wgk %d a_pub %lld prox %lld norm %lld\n", k, wg_klog[0][k] - wg_tlog[0][0], wg_klog[1][k] - wg_tlog[0][0],
                       wg_klog[2][k] - wg_tlog[0][0]);
        }
#endif
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
