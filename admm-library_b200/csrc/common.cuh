// common.cuh -- shared constants and small helpers of libadmm_b200 (device + host).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace admmb {

// per-stage factor record, in doubles (mirrored by the oracle for tests only)
//   K[3][6] Acl[6][6] Hinv[3][4] E[3][6] A[6][6] B[6][4] c[6] chat[6]
// every matrix row starts on a 16-byte boundary (Hinv and B rows are padded to 4 doubles) so that the
// shared-memory copy can be read with 128-bit loads.
constexpr int F_K = 0, F_ACL = 18, F_HINV = 54, F_E = 66, F_A = 84, F_B = 120, F_C = 144,
              F_CHAT = 150, FS = 156;
constexpr int HINV_LD = 4, B_LD = 4;

// packed record of a DECOUPLED model (in-plane states {x,y,vx,vy} = indices {0,1,3,4} with controls
// {0,1}; cross-track states {z,vz} = {2,5} with control {2}): CW, Yamanaka-Ankersen and every other
// linearisation about a Kepler orbit have this structure, and so do all of its Riccati matrices.
// Only the structurally non-zero entries are kept (88 instead of 156 doubles per stage):
//   Kin[2][4] Kc[2] Aclin[4][4] Aclc[2][2] Hin[2][2] Hc,pad Ein[2][4] Ec[2] Ain[4][4] Ac[2][2] Bin[4][2] Bc[2] c[6] chat[6]
constexpr int D_KIN = 0, D_KC = 8, D_ACLIN = 10, D_ACLC = 26, D_HIN = 30, D_HC = 34, D_EIN = 36, D_EC = 44,
              D_AIN = 46, D_AC = 62, D_BIN = 66, D_BC = 74, D_C = 76, D_CHAT = 82, FD = 88;

constexpr int BLK_FREE = 0, BLK_L1 = 1, BLK_L1_BOX = 2, BLK_L2 = 3, BLK_L2_BALL = 4, BLK_BOX = 5,
              BLK_BALL = 6, BLK_POINT = 7, BLK_NONE = 8;
constexpr int PAR_LAM = 0, PAR_RAD = 1, PAR_LO = 2, PAR_HI = 5;
constexpr int ST_CONVERGED = 0, ST_MAX_ITER = 1, ST_NAN = 2, ST_RUNNING = -1;

// kernel variants of the FP64 Riccati path (include/admm_b200.h ADMMB_KERNEL_*)
constexpr int KV_AUTO = 0, KV_THREAD = 1, KV_THREAD_WIDE = 2, KV_THREAD2 = 3, KV_TILE = 4, KV_WG = 5, KV_PINT = 6;

constexpr int NUM_SMS_B200 = 148;
constexpr double RHO_MAX = 1.0e6, RHO_MIN = 1.0e-6;   // adaptive rho never leaves this range

// iterate loads / stores of the fast paths.  The two sweeps touch the rows in opposite orders (the backward
// sweep ends at stage 0, where the forward sweep starts; the forward sweep ends at stage N-1, where the next
// backward sweep starts), so the most recently touched rows are the next ones needed: cache them in L2
// (.cg: L2 only, no L1 allocation) instead of streaming them through with evict-first hints.
#ifndef ADMMB_LD
#define ADMMB_LD __ldcg
#define ADMMB_ST __stcg
#endif

// streaming loads/stores of the iterates: they are touched once per sweep, keep them out of L1
__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double *p, double v)
{
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

}  // namespace admmb
