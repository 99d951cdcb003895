// dense.cuh -- the shared-factor dense x-update (SURVEY 8(a) rows a1' and a2') in FP64:
//   X = M RT + S s0 + mc 1'   as one GEMM over the stacked right-hand sides, [row][ld] layout.
// The FP64 kernel accumulates every output in ascending-k order with explicit fma, i.e. exactly
// the oracle's xupdate_dense order, so it stays bit-comparable.  tcgen05 has no FP64 kind
// (ptxas rejects .kind::f64), so the tensor-core version of this GEMM is the TF32 kernel in
// dense_tf32.cuh; this one runs on the FP64 pipe.
#pragma once
#include "common.cuh"
#include "host_util.cuh"
#include "dense_tf32.cuh"

namespace admmb {

struct DenseState {
    DevBuf<double> M, S, mc;          // [n][n] row-major, [n][6], [n]
    DevBuf<double> x, rt, norms, dscr; // [n][ld], [n][ld], [5][ld], [3N][ld] (Riccati scratch of the condensed path)
    DevBuf<double> bf_rt, bf_s0, bf_d, bf_x;   // scratch of dense_build_factor
    DevBuf<int> running, itbase;      // [1], [1] (iteration counter of the graph replays)
    bool ready = false;
    Tf32Plan tf32;                    // tensor-core operands and TMA descriptors (precision = tf32)
    Tf32Condensed cond;               // the same restricted to the split rows (used when there is no linear cost)
    DevBuf<int> sblk;                 // split block j -> block number
    bool condensed = false;
};

constexpr int DG_BM = 64, DG_BN = 128, DG_BK = 16;

// C[i][p] = mc[i] + sum_j S[i][j] s0[j][p] + sum_l M[i][l] RT[l][p], stored only where status[p] is
// still running (finished problems keep their final x).  256 threads, 64x128 tile, 4x8 per thread.
__global__ void __launch_bounds__(256) k_dense_xupdate_f64(int n, int64_t batch, size_t ld, const double *M,
                                                           const double *S, const double *mc, const double *s0,
                                                           const double *rt, const int *status, double *x)
{
    __shared__ double As[DG_BK][DG_BM + 1];
    __shared__ double Bs[DG_BK][DG_BN];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int i0 = blockIdx.y * DG_BM;
    const int64_t p0 = (int64_t)blockIdx.x * DG_BN;
    double acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t p = p0 + tx + 16 * c;
            double a = 0.0;
            if (i < n && p < batch) {
                a = mc[i];
#pragma unroll
                for (int j = 0; j < 6; ++j) a = fma(S[(size_t)i * 6 + j], s0[(size_t)j * ld + p], a);
            }
            acc[r][c] = a;
        }
    }
    for (int k0 = 0; k0 < n; k0 += DG_BK) {
        // A tile: 64 x 16 (4 per thread), B tile: 16 x 128 (8 per thread)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = tid + 256 * t;
            const int r = e >> 4, kk = e & 15;
            const int i = i0 + r, l = k0 + kk;
            As[kk][r] = (i < n && l < n) ? M[(size_t)i * n + l] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int e = tid + 256 * t;
            const int kk = e >> 7, c = e & 127;
            const int l = k0 + kk;
            const int64_t p = p0 + c;
            Bs[kk][c] = (l < n && p < batch) ? rt[(size_t)l * ld + p] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DG_BK; ++kk) {
            double a[4], b[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = Bs[kk][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t p = p0 + tx + 16 * c;
            if (i < n && p < batch && (!status || status[p] == ST_RUNNING)) x[(size_t)i * ld + p] = acc[r][c];
        }
    }
}

}  // namespace admmb

namespace admmb {

// unit right-hand sides for the dense factor build: column i < n has rt = e_i; column n+j has
// s0 = e_j; column n+6 is all zero (affine term only).
__global__ void k_dense_unit_rhs(int n, size_t ldc, double *rt, double *s0)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rt[(size_t)i * ldc + i] = 1.0;
    else if (i < n + 6) s0[(size_t)(i - n) * ldc + i] = 1.0;
}

// xcols [n][ldc]: columns 0..n-1 -> M (row-major [n][n]), n..n+5 -> S [n][6], n+6 -> mc
__global__ void k_dense_split_factor(int n, size_t ldc, const double *xcols, int has_c, double *M, double *S,
                                     double *mc)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double *row = xcols + (size_t)r * ldc;
    for (int i = 0; i < n; ++i) M[(size_t)r * n + i] = row[i];
    for (int j = 0; j < 6; ++j) S[(size_t)r * 6 + j] = row[n + j];
    mc[r] = has_c ? row[n + 6] : 0.0;
}

}  // namespace admmb
