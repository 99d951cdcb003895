// dense_tf32.cuh -- the shared-factor x-update on the 5th-generation tensor cores (SURVEY 8(a) row a2',
// north_star subsystem (1)): X = [M | S | mc] * [RT; s0; 1] as ONE TF32 GEMM
//     D[i][p] (TMEM, fp32) = sum_k A[i][k] * B[k][p],   A = Mext [Mpad x Kpad] fp32 (K-major),
//                                                       B = RText [Kpad x ld] fp32 (MN-major: problems contiguous)
// written directly against the hardware: TMA (cp.async.bulk.tensor, 128-byte swizzle) feeds a 4-stage
// shared-memory ring, one elected thread issues tcgen05.mma.kind::tf32 with the accumulator in tensor
// memory, tcgen05.commit releases ring slots, four epilogue warps read the accumulator with tcgen05.ld
// and store x.  tcgen05 has no FP64 kind (ptxas: "Unknown modifier '.kind::f64'"), so this path is a
// separate, stated precision class: TF32 products (10-bit mantissa), FP32 accumulation.  With
// split = 3 the operands are split hi + lo and three MMAs per step (Ah*Bh + Ah*Bl + Al*Bh) recover
// roughly FP32 accuracy ("3xTF32") at three times the (cheap) tensor work.
#pragma once
#include <cuda.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "host_util.cuh"

namespace admmb {

constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 32;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;                 // 16 KB
constexpr int TG_B_CHUNK = 32 * TG_BK * 4;                    // one 32-problem-wide MN atom column: 4 KB
constexpr int TG_B_BYTES = (TG_BN / 32) * TG_B_CHUNK;         // 16 KB
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
// ring depth: 4 stages of 32 KB (plain) or 3 stages of 64 KB (hi + lo operands) fit the 227 KB limit
__host__ __device__ constexpr int tg_stages(int split) { return split == 3 ? 3 : 4; }
__host__ __device__ constexpr int tg_smem_bytes(int split) { return 1024 + tg_stages(split) * TG_STAGE_BYTES * (split == 3 ? 2 : 1) + 256; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tg_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TG_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TG_DONE;\n\t"
        "bra TG_WAIT;\n\t"
        "TG_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tg_tma_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor (sm_100 format): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61
// layout 2 = SWIZZLE_128B (K-major A), layout 1 = SWIZZLE_128B_BASE32B: the only layout the hardware accepts for an
// MN-major TF32 operand (128 B along MN x 4 along K atoms, 32-byte swizzle granularity; TMA mode 128B_ATOM_32B)
__device__ __forceinline__ uint64_t tg_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void tg_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tg_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Persistent: grid = min(tiles, SMs); CTA c walks tiles c, c + grid, ... in M-fastest order (the M-tiles that share
// one right-hand-side tile run side by side, so B is fetched from DRAM once and then from L2).  192 threads: warp 0
// TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..5 epilogue (TMEM lane quarter = warp % 4).  The
// accumulator is double-buffered in TMEM (2 x 128 columns): the MMA warp starts tile i+1 while the epilogue warps
// drain tile i, and the TMA ring never drains between tiles.
template <int SPLIT>   // 1: plain TF32; 3: hi/lo split operands (Ah*Bh + Ah*Bl + Al*Bh)
__global__ void __launch_bounds__(192, 1)
k_dense_xupdate_tf32(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                     const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo, int n_rows,
                     int kpad, size_t ldx, int m_tiles, int n_tiles, float *x, float *dbg = nullptr)
{
    extern __shared__ __align__(1024) unsigned char tg_dyn_smem[];
    const uint32_t base = (smem_u32(tg_dyn_smem) + 1023u) & ~1023u;      // 128B swizzle atoms need 1 KB alignment
    constexpr int OPERANDS = SPLIT == 3 ? 2 : 1;
    constexpr int TG_STAGES = tg_stages(SPLIT);
    constexpr uint32_t STAGE = TG_STAGE_BYTES * OPERANDS;
    const uint32_t bars = base + TG_STAGES * STAGE;      // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem slot
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (TG_STAGES + s); };
    auto tmem_full = [&](int b) { return bars + 8u * (2 * TG_STAGES + b); };
    auto tmem_empty = [&](int b) { return bars + 8u * (2 * TG_STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * TG_STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = (kpad + TG_BK - 1) / TG_BK;
    const int tiles = m_tiles * n_tiles;
    (void)dbg;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) { tg_mbar_init(full(s), 1); tg_mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { tg_mbar_init(tmem_full(b), 1); tg_mbar_init(tmem_empty(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            unsigned g = 0;                                             // ring uses so far
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t % m_tiles) * TG_BM, n0 = (t / m_tiles) * TG_BN;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % TG_STAGES;
                    const uint32_t ph = (g / TG_STAGES) & 1u;
                    tg_mbar_wait(empty(s), ph ^ 1u);
                    tg_mbar_expect_tx(full(s), STAGE);
                    const uint32_t sa = base + s * STAGE, sb = sa + TG_A_BYTES * OPERANDS;
                    tg_tma_2d(sa, &mapA, kb * TG_BK, m0, full(s));
                    if (SPLIT == 3) tg_tma_2d(sa + TG_A_BYTES, &mapAlo, kb * TG_BK, m0, full(s));
                    for (int c = 0; c < TG_BN / 32; ++c) {
                        tg_tma_2d(sb + c * TG_B_CHUNK, &mapB, n0 + 32 * c, kb * TG_BK, full(s));
                        if (SPLIT == 3) tg_tma_2d(sb + TG_B_BYTES + c * TG_B_CHUNK, &mapBlo, n0 + 32 * c, kb * TG_BK, full(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, A K-major, B MN-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TG_BN >> 3) << 17) |
                                   ((uint32_t)(TG_BM >> 4) << 24);
            unsigned g = 0, li = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++li) {
                const uint32_t buf = li & 1u;
                tg_mbar_wait(tmem_empty(buf), ((li >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)TG_BN;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % TG_STAGES;
                    const uint32_t ph = (g / TG_STAGES) & 1u;
                    tg_mbar_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = base + s * STAGE, sb = sa + TG_A_BYTES * OPERANDS;
#pragma unroll
                    for (int j = 0; j < TG_BK / 8; ++j) {
                        // A: K-major SW128, one k-step = 32 B inside the 128 B row; 8-row groups 1024 B apart
                        // B: MN-major SW128/32B: one k-step = two 4-row groups (SBO 512 B) = 1024 B; 32-wide MN atoms
                        //    (the four TMA boxes of a stage) are TG_B_CHUNK = 4096 B apart (LBO)
                        const uint64_t ah = tg_desc(sa + 32u * j, 16, 1024, 2);
                        const uint64_t bh = tg_desc(sb + 1024u * j, TG_B_CHUNK, 512, 1);
                        tg_mma_tf32(d_tmem, ah, bh, idesc, (kb | j) != 0);
                        if (SPLIT == 3) {
                            const uint64_t al = tg_desc(sa + TG_A_BYTES + 32u * j, 16, 1024, 2);
                            const uint64_t bl = tg_desc(sb + TG_B_BYTES + 1024u * j, TG_B_CHUNK, 512, 1);
                            tg_mma_tf32(d_tmem, ah, bl, idesc, 1);
                            tg_mma_tf32(d_tmem, al, bh, idesc, 1);
                        }
                    }
                    tg_commit(empty(s));                 // ring slot free once these MMAs have read it
                }
                tg_commit(tmem_full(buf));               // accumulator of this tile complete
            }
        }
    } else {
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        unsigned li = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++li) {
            const int m0 = (t % m_tiles) * TG_BM, n0 = (t / m_tiles) * TG_BN;
            const uint32_t buf = li & 1u;
            tg_mbar_wait(tmem_full(buf), (li >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + 32 * q + lane;
#pragma unroll 1
            for (int c = 0; c < TG_BN / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + buf * (uint32_t)TG_BN + (uint32_t)(32 * c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const size_t col = (size_t)n0 + 32 * c;
                if (row < n_rows && col < ldx) {
                    float4 *dst = reinterpret_cast<float4 *>(x + (size_t)row * ldx + col);
#pragma unroll
                    for (int v = 0; v < 8; ++v)
                        dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                             __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                }
            }
            // this warp is done with the accumulator buffer: let the MMA warp reuse it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(buf)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// ---- operand preparation ---------------------------------------------------------------------------
// Mext [Mpad][Kpad] fp32 (+ low parts): columns 0..n-1 = M, n..n+5 = S, n+6 = mc, the rest zero
__global__ void k_tf32_pack_factor(int n, int mpad, int kpad, const double *M, const double *S, const double *mc,
                                   float *hi, float *lo)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (k >= kpad || i >= mpad) return;
    double v = 0.0;
    if (i < n) v = k < n ? M[(size_t)i * n + k] : (k < n + 6 ? S[(size_t)i * 6 + (k - n)] : (k == n + 6 ? mc[i] : 0.0));
    // hi = the TF32-representable head (10 explicit mantissa bits), lo = what fp32 still holds of the rest
    const float f = (float)v;
    const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
    hi[(size_t)i * kpad + k] = h;
    if (lo) lo[(size_t)i * kpad + k] = (float)(v - (double)h);
}

// RText rows n..n+5 = s0, row n+6 = 1, rows n+7.. = 0 (rows 0..n-1 are written by the prox kernel each iteration)
__global__ void k_tf32_pack_tail(int n, int kpad, int64_t batch, size_t ld, const double *s0, float *hi, float *lo)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)ld) return;
    for (int k = n; k < kpad; ++k) {
        double v = 0.0;
        if (p < batch) v = k < n + 6 ? s0[(size_t)(k - n) * ld + p] : (k == n + 6 ? 1.0 : 0.0);
        const float f = (float)v;
        const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
        hi[(size_t)k * ld + p] = h;
        if (lo) lo[(size_t)k * ld + p] = (float)(v - (double)h);
    }
}

// double [rows][ld] -> hi/lo fp32 (unit entry point: converts a caller-supplied RT)
__global__ void k_tf32_split_rows(int rows, size_t ld, const double *in, float *hi, float *lo)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (p >= ld || r >= rows) return;
    const double v = in[(size_t)r * ld + p];
    const float f = (float)v;
    const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
    hi[(size_t)r * ld + p] = h;
    if (lo) lo[(size_t)r * ld + p] = (float)(v - (double)h);
}

__global__ void k_tf32_to_double(int rows, size_t ld, const float *in, double *out)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (p < ld && r < rows) out[(size_t)r * ld + p] = (double)in[(size_t)r * ld + p];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw CudaFail{cudaErrorNotSupported, "cuTensorMapEncodeTiled unavailable"};
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D fp32 tensor [rows][cols] (cols contiguous, row pitch `pitch` elements), box = box_cols x box_rows, 128B swizzle
inline CUtensorMap make_map_2d(const float *ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_cols, uint32_t box_rows,
                               CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaFail{cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed"};
    return m;
}

// one GEMM launch: X[0..n_rows) = A (mpad x kpad) * B (kpad x ld)
inline void tf32_launch_gemm(int split, const CUtensorMap &mA, const CUtensorMap &mAlo, const CUtensorMap &mB,
                             const CUtensorMap &mBlo, int n_rows, int mpad, int kpad, size_t ld, float *X, cudaStream_t st,
                             float *dbg = nullptr)
{
    const int m_tiles = mpad / TG_BM, n_tiles = (int)((ld + TG_BN - 1) / TG_BN);
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        sms = NUM_SMS_B200;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const unsigned grid = (unsigned)std::min(m_tiles * n_tiles, sms);
    static bool attr_done[64] = {};   // the attribute is per device; set it once, not per launch
    int cur = 0;
    CK(cudaGetDevice(&cur));
    bool &attr_set = attr_done[cur & 63];
    if (!attr_set) {
        CK(cudaFuncSetAttribute(k_dense_xupdate_tf32<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem_bytes(3)));
        CK(cudaFuncSetAttribute(k_dense_xupdate_tf32<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem_bytes(1)));
        attr_set = true;
    }
    if (split == 3)
        k_dense_xupdate_tf32<3><<<grid, 192, tg_smem_bytes(3), st>>>(mA, mAlo, mB, mBlo, n_rows, kpad, ld, m_tiles, n_tiles, X, dbg);
    else
        k_dense_xupdate_tf32<1><<<grid, 192, tg_smem_bytes(1), st>>>(mA, mAlo, mB, mBlo, n_rows, kpad, ld, m_tiles, n_tiles, X, dbg);
}

struct Tf32Plan {
    int n = 0, mpad = 0, kpad = 0, split = 1;
    size_t ld = 0;
    DevBuf<float> Ahi, Alo, Bhi, Blo, X;
    CUtensorMap mA, mAlo, mB, mBlo;
    void prepare(int n_, int64_t batch, size_t ld_, int split_, const double *M, const double *S, const double *mc,
                 const double *s0, cudaStream_t st)
    {
        n = n_; ld = ld_; split = split_;
        mpad = (int)round_up((size_t)n, TG_BM);
        kpad = (int)round_up((size_t)n + 7, 8);
        Ahi.alloc((size_t)mpad * kpad);
        Bhi.alloc((size_t)kpad * ld);
        X.alloc((size_t)mpad * ld);
        if (split == 3) { Alo.alloc((size_t)mpad * kpad); Blo.alloc((size_t)kpad * ld); }
        dim3 g1((unsigned)((kpad + 127) / 128), (unsigned)mpad);
        k_tf32_pack_factor<<<g1, 128, 0, st>>>(n, mpad, kpad, M, S, mc, Ahi.p, split == 3 ? Alo.p : nullptr);
        CK(cudaMemsetAsync(Bhi.p, 0, sizeof(float) * kpad * ld, st));
        if (split == 3) CK(cudaMemsetAsync(Blo.p, 0, sizeof(float) * kpad * ld, st));
        k_tf32_pack_tail<<<(unsigned)((ld + 127) / 128), 128, 0, st>>>(n, kpad, batch, ld, s0, Bhi.p, split == 3 ? Blo.p : nullptr);
        CK(cudaGetLastError());
        mA = make_map_2d(Ahi.p, mpad, kpad, kpad, TG_BK, TG_BM);
        mB = make_map_2d(Bhi.p, kpad, ld, ld, 32, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        mAlo = split == 3 ? make_map_2d(Alo.p, mpad, kpad, kpad, TG_BK, TG_BM) : mA;
        mBlo = split == 3 ? make_map_2d(Blo.p, kpad, ld, ld, 32, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : mB;
    }
    float *dbg = nullptr;
    void gemm(cudaStream_t st) { tf32_launch_gemm(split, mA, mAlo, mB, mBlo, n, mpad, kpad, ld, X.p, st, dbg); }
};

// ---- condensed incremental form ----------------------------------------------------------------------
// When the states are unsplit (BLK_NONE), every control block is split and there is no linear cost, the
// right-hand side is zero outside the split rows R and the prox only reads x on R, so the per-iteration GEMM
// shrinks to the R x R block of M (CW, N = 50: 153 instead of 456 along both M and K: 9x fewer flops).
// The x-update is affine in the right-hand side, so the GEMM is applied to its INCREMENT:
//     x_R^{k+1} = x_R^k + M_RR (rt^k - rt^{k-1}),      x_R^1 from one exact FP64 Riccati x-update,
// with x_R accumulated in FP64 by the prox kernel.  The TF32/FP32 rounding of the tensor-core product is then
// relative to the size of the step, which shrinks as ADMM converges, instead of relative to x itself -- the
// absolute form stalls at |r| ~ 1e-6 |x| and cannot meet a 1e-6 tolerance; the incremental form can.
// The accumulated x_R only drives the iteration.  When a problem finishes at iteration k, the x it returns is the
// EXACT FP64 Riccati x-update of the right-hand side its last iteration used, rt^{k-1} = (z^k - u^k) - increment^k
// (the increment is still in the buffer: a finished problem's column is never written again), so the returned
// trajectory satisfies the dynamics to FP64 round-off and carries no accumulated tensor-core rounding
// (k_tf32_final_x in kernels.cuh).

// A[i][k] = M[rmap[i]][cmap[k]], zero padded
__global__ void k_tf32_pack_factor_cond(int n, int nr, int mpad, int kpad, const int *rows, const double *M, float *hi,
                                        float *lo)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (k >= kpad || i >= mpad) return;
    const double v = (i < nr && k < nr) ? M[(size_t)rows[i] * n + rows[k]] : 0.0;
    const float f = (float)v;
    const float h = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
    hi[(size_t)i * kpad + k] = h;
    if (lo) lo[(size_t)i * kpad + k] = (float)(v - (double)h);
}

// xacc[i][p] = x[rows[i]][p]
__global__ void k_tf32_gather_rows(int nr, const int *rows, int64_t batch, size_t ld, const double *x, double *xacc)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (p < batch && i < nr) xacc[(size_t)i * ld + p] = x[(size_t)rows[i] * ld + p];
}

// increment = 0 for the still-running columns (after a refresh of x_R)
__global__ void k_tf32_zero_running(int rows, int64_t width, size_t ld, const int *status, float *hi, float *lo)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (p >= width || r >= rows || status[p] != ST_RUNNING) return;
    hi[(size_t)r * ld + p] = 0.f;
    if (lo) lo[(size_t)r * ld + p] = 0.f;
}

struct Tf32Condensed {
    int n = 0, nr = 0, mpad_r = 0, kpad = 0, split = 1;
    size_t ld = 0, ld_cur = 0;
    DevBuf<float> Ahi, Alo;          // M_RR
    DevBuf<float> Bhi, Blo;          // right-hand-side increment (home copy), [kpad][ld]
    DevBuf<float> Xr;                // M_RR * increment, [mpad_r][ld]
    DevBuf<double> Xacc;             // accumulated x_R (home copy), [nr][ld]
    DevBuf<int> rows;                // R
    float *bh = nullptr, *bl = nullptr;   // the increment buffers of the current working set
    CUtensorMap mA, mAlo, mB, mBlo;
    void prepare(int n_, const std::vector<int> &R, size_t ld_, int split_, const double *M, cudaStream_t st)
    {
        n = n_; nr = (int)R.size(); ld = ld_; split = split_;
        mpad_r = (int)round_up((size_t)nr, TG_BM);
        kpad = (int)round_up((size_t)nr, 8);
        const bool s3 = split == 3;
        rows.alloc(nr);
        CK(cudaMemcpyAsync(rows.p, R.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, st));
        Ahi.alloc((size_t)mpad_r * kpad);
        if (s3) Alo.alloc((size_t)mpad_r * kpad);
        dim3 g1((unsigned)((kpad + 127) / 128), (unsigned)mpad_r);
        k_tf32_pack_factor_cond<<<g1, 128, 0, st>>>(n, nr, mpad_r, kpad, rows.p, M, Ahi.p, s3 ? Alo.p : nullptr);
        CK(cudaGetLastError());
        Bhi.alloc((size_t)kpad * ld);
        CK(cudaMemsetAsync(Bhi.p, 0, sizeof(float) * kpad * ld, st));       // first increment is zero: x^1 is exact
        if (s3) { Blo.alloc((size_t)kpad * ld); CK(cudaMemsetAsync(Blo.p, 0, sizeof(float) * kpad * ld, st)); }
        Xr.alloc((size_t)mpad_r * ld);
        Xacc.alloc((size_t)nr * ld);
        mA = make_map_2d(Ahi.p, mpad_r, kpad, kpad, TG_BK, TG_BM);
        mAlo = s3 ? make_map_2d(Alo.p, mpad_r, kpad, kpad, TG_BK, TG_BM) : mA;
        bind(Bhi.p, Blo.p, ld);
    }
    // point the GEMM at the increment buffers of the current working set (pitch ld_now)
    void bind(float *h, float *l, size_t ld_now)
    {
        const bool s3 = split == 3;
        ld_cur = ld_now;
        bh = h; bl = s3 ? l : nullptr;
        mB = make_map_2d(bh, kpad, ld_cur, ld_cur, 32, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        mBlo = s3 ? make_map_2d(bl, kpad, ld_cur, ld_cur, 32, TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : mB;
    }
    void gemm(cudaStream_t st) { tf32_launch_gemm(split, mA, mAlo, mB, mBlo, nr, mpad_r, kpad, ld_cur, Xr.p, st); }
};

}  // namespace admmb
