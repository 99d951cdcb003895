/* admm_b200.h -- C ABI of libadmm_b200.so, the B200-native batched ADMM solver for convex
 * optimal-control QPs (fuel-optimal rendezvous, impulsive / low-thrust transfers on linearised
 * CW / Yamanaka-Ankersen / STM-propagated dynamics).
 *
 * Reference interface replaced: NONE EXISTS.  /root/reference/README.md:1-2 is the whole of the
 * reference's non-licence content ("Implementation of Alternating Direction Method of Multipliers
 * for astrodynamics problems"); it ships no function, FFI or plugin surface.  The kind of boundary
 * is fixed by BASELINE.json `north_star`: "MATLAB host code calls CUDA through a thin C-ABI MEX
 * layer ... problem struct in, primal/dual iterates and residual history out, same rho/alpha/
 * tolerance semantics".  Each entry point below names the row of SURVEY.md section 8 it serves and
 * the oracle function (oracle/admm_ocp.m, the MATLAB text the north_star mandates) it replaces.
 *
 * Everything is plain C: POD structs of sizes, flags and raw HOST pointers.  All arrays use MATLAB
 * column-major layout so that the MEX gateway (admm-library_b200/mex/admm_mex.cpp) is zero-copy:
 *
 *   x = (s_0,a_0,...,s_{N-1},a_{N-1},s_N), n = 9N+6, 6 states [r;v] and 3 controls per stage;
 *   every consecutive 3-vector of x is one prox block, nb = 3N+2.
 *   A [6x6xNxBd]  B [6x3xNxBd]  c [6xNxBd]  Q [6x6x(N+1)xBd]  R [3x3xNxBd]   (Bd = 1 or batch)
 *   q [n x Bq]  s0 [6 x batch]  block_type [nb] int32  block_par [8 x nb x Bp]
 *   z0,u0 [n x batch]  rho0 [batch]     outputs x,z,u [n x batch], history [max_iter x batch].
 *
 * Ownership: the caller owns every input and output buffer; the library never writes an input
 * and never frees anything it did not allocate.  Device memory, streams and worker threads belong
 * to the handle.  One call in flight per handle.  No CPU fallback: without a usable CUDA device
 * admmb_create returns ADMMB_E_NODEVICE.
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMMB_VERSION 210

/* return codes */
enum {
    ADMMB_OK = 0,
    ADMMB_E_BADARG = -1,
    ADMMB_E_CUDA = -2,
    ADMMB_E_NCCL = -3,      /* a call of the statistics all-reduce failed (multi-GPU handles only) */
    ADMMB_E_NOMEM = -4,
    ADMMB_E_NODEVICE = -5,
    ADMMB_E_STATE = -6      /* staged API called out of order */
};

/* prox block types (oracle/admm_ocp.m prox_blocks) */
enum {
    ADMMB_BLK_FREE = 0,     /* g = 0, block is split (x_b = z_b enforced)                      */
    ADMMB_BLK_L1 = 1,       /* lam |v|_1                                                       */
    ADMMB_BLK_L1_BOX = 2,   /* lam |v|_1 + indicator(lo <= v <= hi)                            */
    ADMMB_BLK_L2 = 3,       /* lam |v|_2                                                       */
    ADMMB_BLK_L2_BALL = 4,  /* lam |v|_2 + indicator(|v|_2 <= rad)  (thrust-magnitude SOC)     */
    ADMMB_BLK_BOX = 5,      /* indicator(lo <= v <= hi)                                        */
    ADMMB_BLK_BALL = 6,     /* indicator(|v - cen|_2 <= rad)                                   */
    ADMMB_BLK_POINT = 7,    /* indicator(v == cen)         (terminal-set projection)           */
    ADMMB_BLK_NONE = 8      /* block takes no part in the splitting (no z, no u)               */
};
/* block_par slots: [0] lam, [1] rad, [2..4] lo[3] or cen[3], [5..7] hi[3] */

/* per-problem status */
enum { ADMMB_ST_CONVERGED = 0, ADMMB_ST_MAX_ITER = 1, ADMMB_ST_NAN = 2 };

/* x-update selection (SURVEY 8(a) rows a2 / a2') and arithmetic */
enum { ADMMB_XUPDATE_AUTO = 0, ADMMB_XUPDATE_DENSE = 1, ADMMB_XUPDATE_RICCATI = 2 };
enum { ADMMB_PREC_FP64 = 0, ADMMB_PREC_TF32 = 1 };   /* TF32: tensor cores allowed -- with XUPDATE_DENSE throughout, with
                                                        XUPDATE_AUTO once the working set is narrow; not with RICCATI */

/* kernel selection of the FP64 Riccati path.  AUTO in normal use (the library picks by working-set width); the other
 * codes pin one variant so that tests can hold EVERY variant against the oracle at any batch size.  A variant that
 * is not applicable to the problem (e.g. TILE with per-problem models) falls back to AUTO's choice. */
enum {
    ADMMB_KERNEL_AUTO = 0,
    ADMMB_KERNEL_THREAD = 1,       /* one problem per thread, 128-register build, two CTAs per SM            */
    ADMMB_KERNEL_THREAD_WIDE = 2,  /* one problem per thread, uncapped registers, deeper prefetch            */
    ADMMB_KERNEL_THREAD2 = 3,      /* two problems per thread (full-width working sets)                      */
    ADMMB_KERNEL_TILE = 4,         /* resident tile: one warp per 32 problems, iterates in shared memory     */
    ADMMB_KERNEL_WG = 5,           /* warp group: eight role-split warps per 32-problem resident tile        */
    ADMMB_KERNEL_PINT = 6          /* parallel in time: eight warps sweep eight chunks of stages at the
                                      same time, chunk boundaries by superposition.  FP64, but not the oracle's operation
                                      order: x, z, u agree with it to ~1e-12 relative, iteration counts to +-1 on threshold
                                      cases.  OPT-IN only: AUTO never picks it (the default path is bit-exact)          */
};

typedef struct admmb_ctx *admmb_handle;

/* `prob` of [x,z,u,hist] = admm_solve(prob, opts)  (oracle/admm_ocp.m) */
typedef struct admmb_problem {
    int32_t N;                   /* horizon: number of stages                                  */
    int64_t batch;               /* number of independent problems                              */
    const double *A;             /* [6x6xNxBd]                                                  */
    int32_t dyn_batched;         /* 0: A,B,c,Q,R shared by the batch; 1: one model per problem  */
    const double *B;             /* [6x3xNxBd]                                                  */
    const double *c;             /* [6xNxBd]            or NULL                                 */
    const double *Q;             /* [6x6x(N+1)xBd]      or NULL  (state cost, PSD)              */
    const double *R;             /* [3x3xNxBd]          or NULL  (control cost, PSD)            */
    const double *q;             /* [n x Bq]            or NULL  (linear cost)                  */
    int32_t q_batched;
    const double *s0;            /* [6 x batch] initial states                                  */
    const int32_t *block_type;   /* [nb] ADMMB_BLK_*, shared by the batch                       */
    const double *block_par;     /* [8 x nb x Bp]                                               */
    int32_t par_batched;
    const double *z0, *u0;       /* [n x batch] warm start or NULL (zeros)                      */
    const double *rho0;          /* [batch] per-problem initial rho or NULL (opts.rho)          */
} admmb_problem;

/* `opts`.  rho / alpha / abstol / reltol have Boyd et al. (2011) semantics. */
typedef struct admmb_opts {
    double rho, alpha, abstol, reltol;
    int32_t max_iter;
    int32_t adapt_rho;           /* residual balancing on/off                                   */
    double adapt_mu, adapt_tau;
    int32_t adapt_every;         /* test every this many iterations                             */
    int32_t adapt_until;         /* no adaptation after this iteration (0 = no limit)           */
    int32_t xupdate;             /* ADMMB_XUPDATE_*                                             */
    int32_t precision;           /* ADMMB_PREC_*                                                */
    int32_t history;             /* record per-iteration r,s,eps,rho                            */
    int32_t chunk;               /* iterations per persistent launch (0 = library default)     */
    int32_t kernel;              /* ADMMB_KERNEL_* (0 = auto)                                   */
    int32_t tf32_switch;         /* precision = TF32, xupdate = auto: running problems at which the tensor-core
                                    pair takes over from the FP64 Riccati kernel (0 = default 8192; < 0 = never) */
    int32_t tf32_refresh;        /* TF32 path: iterations between exact FP64 refreshes of x (0 = default 1000;
                                    < 0 = never)                                                 */
} admmb_opts;

/* outputs; every pointer is caller-allocated and may be NULL when that output is not wanted */
typedef struct admmb_result {
    double *x, *z, *u;                         /* [n x batch]                                   */
    int32_t *iters, *status;                   /* [batch]                                       */
    double *r_norm, *s_norm, *eps_pri, *eps_dual, *rho;          /* [batch] finals              */
    double *hist_r, *hist_s, *hist_eps_pri, *hist_eps_dual, *hist_rho;   /* [max_iter x batch]  */
    int64_t stats[4];            /* converged count, sum of iterations, max iterations, refactors */
    double device_ms;            /* device time of the solve phase (CUDA events, max over GPUs) */
    double h2d_ms, d2h_ms;       /* upload (+layout +factor) and download phases                */
    int64_t launches;            /* kernels launched by this library during the call            */
    double kernel_ms;            /* summed device time of the dominant kernel's launches (the
                                    persistent ADMM iteration kernel, or GEMM + prox on the dense
                                    path), CUDA events on the launching stream, max over GPUs    */
    int64_t kernel_launches;     /* how many launches kernel_ms covers                          */
} admmb_result;

/* ---- lifetime ---------------------------------------------------------------------------- */
int admmb_version(void);
/* device_ids == NULL: use devices 0..n_devices-1; n_devices <= 0: all visible devices. */
int admmb_create(admmb_handle *out, const int *device_ids, int n_devices);
int admmb_destroy(admmb_handle h);
const char *admmb_last_error(admmb_handle h);          /* h may be NULL: last create() error   */
int admmb_device_count(admmb_handle h);
/* how many statistics all-reduces (result.stats of a multi-GPU handle, 4 + 1 integers per GPU over NCCL /
 * NVLink, libnccl opened at run time) this handle has done; 0: one GPU, or libnccl not found (host-side sum) */
int admmb_nccl_gathers(admmb_handle h);

/* ---- the solve (admm_solve) --------------------------------------------- */
/* upload + run + download in one blocking call; shards `batch` contiguously over the handle's GPUs. */
int admmb_solve(admmb_handle h, const admmb_problem *prob, const admmb_opts *opts, admmb_result *res);

/* staged form, for callers that keep the batch resident in HBM (benchmarks, MPC loops):
 *   upload  : H2D of the problem, AoS->SoA layout change on device, factorisation (rows a1/a1')
 *   run     : resets the iterates to the uploaded warm start and iterates to tolerance on device
 *   download: layout change back and D2H of whichever outputs are non-NULL                      */
int admmb_upload(admmb_handle h, const admmb_problem *prob, const admmb_opts *opts);
int admmb_run(admmb_handle h, const admmb_opts *opts, admmb_result *stats_only);
int admmb_download(admmb_handle h, admmb_result *res);
/* receding-horizon step on the resident batch: the (z, u) of the solve just done, shifted by k stages
 * (stage j starts from stage j+k; the last k stages from zero; terminal blocks keep theirs), become the warm start,
 * the initial states become s0_new [6 x batch] (NULL: the solution's own state at stage k), and the batch is solved
 * again without any re-upload of the model.  Riccati path only.  Equivalent to admmb_solve with z0 / u0 / s0 set so. */
int admmb_shift_resolve(admmb_handle h, int32_t k, const double *s0_new, const admmb_opts *opts,
                        admmb_result *stats_only);
/* ---- on-device problem generators ----------------------------------------------------------------
 * The stage matrices A_k, B_k are computed on the GPU from a few parameters instead of being uploaded: a per-problem
 * model is 54 N doubles per problem over PCIe (354 MB for 16,384 problems at N = 50), its parameters are 16 B.
 * Oracle: oracle/gen_ocp.py (same IEEE operations in the same order; outputs are bit-identical). */
enum {
    ADMMB_GEN_CW_IMPULSIVE = 1,  /* shared model: A = Phi_CW(T), B = Phi_CW(T)[:, 3:6]            (closed form)     */
    ADMMB_GEN_CW_ZOH = 2,        /* shared model: A = Phi_CW(T), B = int_0^T Phi_CW [0; I]        (closed form)     */
    ADMMB_GEN_ELLIPTIC_ZOH = 3   /* per-problem, time-varying: RK4 of the LVLH dynamics linearised about a Kepler
                                    orbit of eccentricity e[p] from true anomaly theta0[p], zero-order-hold input    */
};
typedef struct admmb_generator {
    int32_t kind;                /* ADMMB_GEN_*                                                  */
    int32_t substeps;            /* ELLIPTIC: RK4 steps per stage (0 = 8)                       */
    double T;                    /* stage length (time unit 1 / mean motion)                    */
    double nmm;                  /* CW kinds: mean motion (0 = 1)                               */
    const double *e;             /* ELLIPTIC: [batch] eccentricities, 0 <= e < 1                */
    const double *theta0;        /* ELLIPTIC: [batch] true anomaly at the start of stage 0      */
} admmb_generator;
/* admmb_upload / admmb_solve with prob->A, prob->B and prob->dyn_batched ignored (A, B may be NULL): the model comes
 * from `gen`; c, Q, R must then be shared (or NULL) for the CW kinds and per-problem (or NULL) for ELLIPTIC. */
int admmb_upload_generated(admmb_handle h, const admmb_problem *prob, const admmb_generator *gen,
                           const admmb_opts *opts);
int admmb_solve_generated(admmb_handle h, const admmb_problem *prob, const admmb_generator *gen,
                          const admmb_opts *opts, admmb_result *res);
/* ---- sequential convex programming on the resident batch ------------------------------------------
 * The caller AFTER the hot path: nonlinear dynamics are re-linearised about each problem's own reference trajectory ON
 * THE DEVICE (RK4 of the state and its variational equations per stage -> per-problem A_k, B_k, c_k written straight
 * into the arrays the factor kernel reads), the batched ADMM kernels solve the convex subproblem warm-started from the
 * previous pass, and a problem leaves the loop when max|x - x_ref| <= tol_abs + tol_rel max|x| and its convex solve
 * converged (per-problem early exit: results do not depend on how the batch is sharded; opts->max_iter may therefore cap
 * the early passes -- inexact SCP).  The first reference is the free drift from s0.  Only s0, the block
 * table and (optionally) q / per-problem Q, R are uploaded; prob->A, B, c, dyn_batched are ignored.
 * Oracle: oracle/scp_ocp.py (same IEEE operations in the same order; every pass is bit-identical). */
enum {
    ADMMB_SCP_NL_CIRCULAR = 1,   /* deputy about a chief on a circular orbit of radius R0: LVLH frame, full two-body
                                    gravity, zero-order-hold thrust acceleration (linearised at r = 0: Clohessy-Wiltshire) */
    ADMMB_SCP_NL_ELLIPTIC = 2    /* chief on a Kepler orbit of eccentricity e[p] from true anomaly theta0[p]; R0 is the
                                    semi-major axis, the time unit 1 / mean motion (nmm must be 0 or 1); linearised at
                                    r = 0: the model of ADMMB_GEN_ELLIPTIC_ZOH (config 4)                               */
};
enum {
    ADMMB_SCP_CTRL_ZOH = 0,       /* thrust acceleration held over the stage (low-thrust transfers, configs 3 / 4)   */
    ADMMB_SCP_CTRL_IMPULSIVE = 1  /* velocity increment at the start of the stage, then a coast (configs 1, 2, 5)    */
};
typedef struct admmb_scp {
    int32_t model;               /* ADMMB_SCP_*                                                  */
    int32_t substeps;            /* RK4 steps per stage (0 = 8)                                  */
    double T;                    /* stage length                                                 */
    double nmm;                  /* mean motion (0 = 1)                                          */
    double R0;                   /* radius of the chief's orbit, in the problem's length unit    */
    int32_t max_pass;            /* linearise + solve passes at most (>= 1)                      */
    double tol_abs, tol_rel;     /* per-problem stop: max|x - x_ref| <= tol_abs + tol_rel max|x|, last solve converged */
    int32_t control;             /* ADMMB_SCP_CTRL_*: how the three controls of a stage act                          */
    const double *e;             /* NL_ELLIPTIC: [batch] eccentricities, 0 <= e < 1                                   */
    const double *theta0;        /* NL_ELLIPTIC: [batch] true anomaly at the start of stage 0                         */
} admmb_scp;
/* per-problem SCP outputs; caller-allocated, any pointer may be NULL */
typedef struct admmb_scp_result {
    int32_t *passes;             /* [batch] passes made                                          */
    int32_t *scp_status;         /* [batch] 0: trajectory converged, 1: max_pass reached         */
    double *step;                /* [batch] max|x - x_ref| of the last pass                      */
    int64_t *iters_total;        /* [batch] ADMM iterations over all passes                      */
    double *hist_step;           /* [max_pass x batch] step per pass, NaN after the exit         */
    int64_t stats[4];            /* SCP-converged problems, sum of ADMM iterations over all passes, max passes,
                                    sum of passes                                                 */
    double linearise_ms;         /* device time of the linearisation kernels (CUDA events, max over GPUs) */
} admmb_scp_result;
/* upload + SCP loop + download.  `res` receives x, z, u, iters, status and the finals of every problem's LAST convex
 * solve; res->stats[1] counts the ADMM iterations of all passes, res->device_ms the whole loop.  opts: adapt_rho and
 * history must be off, xupdate auto / riccati.  */
int admmb_scp_solve(admmb_handle h, const admmb_problem *prob, const admmb_scp *scp, const admmb_opts *opts,
                    admmb_result *res, admmb_scp_result *scp_res);
/* launch the device work of GPU 0 on a caller-owned cudaStream_t (NULL: the library's own) */
int admmb_set_stream(admmb_handle h, void *cuda_stream);

/* ---- unit entry points: one kernel each, host buffers in / out (tests, SURVEY 4.2 tier T2) ---- */
/* row a1: shared-model Riccati factor; fac_out [156 x N] records (layout in DESIGN.md)           */
int admmb_k_riccati_factor(admmb_handle h, int32_t N, const double *A, const double *B, const double *c,
                           const double *Q, const double *R, double rho, const int32_t *block_type,
                           double *fac_out);
/* row a2: X = riccati x-update of RT [n x batch] with a shared factor                            */
int admmb_k_xupdate_riccati(admmb_handle h, int32_t N, int64_t batch, const double *fac, int32_t has_c,
                            const double *s0, const double *rt, double *x);
/* rows a3+a4: in-place z,u update and the five squared norms [5 x batch] for given x             */
int admmb_k_prox_dual_residuals(admmb_handle h, int32_t N, int64_t batch, const int32_t *block_type,
                                const double *block_par, int32_t par_batched, const double *rinv,
                                double alpha, const double *x, double *z, double *u, double *norms);
/* row a1': dense shared factor M [n x n] (row-major), S [n x 6], mc [n] from a Riccati factor    */
int admmb_k_dense_factor(admmb_handle h, int32_t N, const double *fac, int32_t has_c,
                         double *M, double *S, double *mc);
/* row a2': X = M RT + S s0 + mc; precision = ADMMB_PREC_FP64 (bit-exact order) or ADMMB_PREC_TF32 */
int admmb_k_xupdate_dense(admmb_handle h, int32_t N, int64_t batch, const double *M, const double *S,
                          const double *mc, const double *s0, const double *rt, int32_t precision,
                          double *x);

/* the generated stage matrices themselves, A [6x6xNxBd], B [6x3xNxBd] (Bd = 1 for the CW kinds, batch
 * for ELLIPTIC), for comparison with oracle/gen_ocp.py                                              */
int admmb_k_generate(admmb_handle h, int32_t N, int64_t batch, const admmb_generator *gen, double *A, double *B);

/* the stage records of one linearisation pass about xref [n x batch] (shoot = 0), or along the nonlinear
 * trajectory from s0 [6 x batch] under the controls of xref (shoot = 1; xref is then overwritten with that trajectory):
 * A [6x6xNxbatch], B [6x3xNxbatch], c [6xNxbatch], for comparison with oracle/scp_ocp.py linearise / shoot       */
int admmb_k_scp_linearise(admmb_handle h, int32_t N, int64_t batch, const admmb_scp *scp, int32_t shoot,
                          const double *s0, double *xref, double *A, double *B, double *c);

#ifdef __cplusplus
}
#endif
#endif /* ADMM_B200_H */
