/* ORACLE (test infrastructure, NOT product code) -- C restatement of oracle/admm_ocp.m with the
 * CANONICAL OPERATION ORDER.
 *
 * PARITY STATUS: unpinned by the reference.  /root/reference/ holds README.md:1-2 and
 * LICENSE:1-21 only; BASELINE.json `north_star` mandates "a minimal MATLAB ADMM written to the
 * README's stated algorithm" as oracle + CPU baseline.  oracle/admm_ocp.m is that text,
 * oracle/admm_ocp.py its NumPy restatement, this file the compiled restatement used
 *   (1) as the bit-level checker of the CUDA path: every floating-point operation below is
 *       written explicitly (mul / add / fma, fixed summation order) and the file is compiled
 *       with -ffp-contract=off, so a CUDA kernel that performs the same operations in the same
 *       order (compiled with -fmad=false, explicit fma()) reproduces x, z, u, the residual
 *       history and therefore the iteration counts BIT FOR BIT;
 *   (2) as the CPU throughput baseline (OpenMP over problems), bench.py `cpu_baseline` and
 *       `--impl reference`.
 * Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load this library.
 *
 * Functions follow SURVEY.md section 8(a): a1 riccati_factor, a1' kkt_dense_factor (built from
 * unit-vector Riccati solves), a2 xupdate_riccati, a2' xupdate_dense, a3 prox_block,
 * a4 dual_and_residuals (fused with a3 per block, same per-accumulator order), a5 adapt_rho,
 * a6 ocp_solve.
 *
 * Host array layout = MATLAB column-major, exactly the C-ABI layout of include/admm_b200.h:
 *   A [6x6xNxBd]  B [6x3xNxBd]  c [6xNxBd]  Q [6x6x(N+1)xBd]  R [3x3xNxBd]  q [n x Bq]
 *   s0 [6 x Bsz]  block_type [nb]  block_par [8 x nb x Bp]  z0,u0 [n x Bsz]  rho0 [Bsz]
 *   outputs x,z,u [n x Bsz]; history [max_iter x Bsz].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { BLK_FREE = 0, BLK_L1, BLK_L1_BOX, BLK_L2, BLK_L2_BALL, BLK_BOX, BLK_BALL, BLK_POINT, BLK_NONE };
enum { PAR_LAM = 0, PAR_RAD = 1, PAR_LO = 2, PAR_HI = 5 };
enum { ST_CONVERGED = 0, ST_MAX_ITER = 1, ST_NAN = 2 };

/* per-stage factor record (doubles) */
enum { F_K = 0, F_ACL = 18, F_HINV = 54, F_E = 63, F_A = 81, F_B = 117, F_C = 135, F_CHAT = 141, FS = 148 };

typedef struct {
    int32_t N;
    int64_t batch;
    const double *A; int32_t dyn_batched;     /* A,B,c,Q,R share one flag: 0 shared, 1 per problem */
    const double *B;
    const double *c;                           /* may be NULL */
    const double *Q;                           /* may be NULL */
    const double *R;                           /* may be NULL */
    const double *q; int32_t q_batched;        /* may be NULL */
    const double *s0;
    const int32_t *block_type;
    const double *block_par; int32_t par_batched;
    const double *z0, *u0, *rho0;              /* may be NULL */
} ocp_problem;

typedef struct {
    double rho, alpha, abstol, reltol;
    int32_t max_iter;
    int32_t adapt_rho; double adapt_mu, adapt_tau; int32_t adapt_every;
    int32_t adapt_until;                       /* no adaptation after this iteration (0 = no limit) */
    int32_t xupdate;                           /* 0 auto(riccati), 1 dense, 2 riccati */
    int32_t history;
} ocp_opts;

typedef struct {
    double *x, *z, *u;                         /* [n x Bsz], may be NULL */
    int32_t *iters, *status;                   /* [Bsz] */
    double *r_norm, *s_norm, *eps_pri, *eps_dual, *rho;   /* [Bsz] finals */
    double *hist_r, *hist_s, *hist_eps_pri, *hist_eps_dual, *hist_rho; /* [max_iter x Bsz] or NULL */
    int64_t stats[4];                          /* converged, sum iters, max iters, refactor count */
    double seconds;                            /* wall time of the solve loop */
} ocp_result;

/* ------------------------------------------------------------------ 3x3 Cholesky helpers */
static int chol3(const double H[3][3], double L[6])
{   /* L = [l00 l10 l11 l20 l21 l22], lower triangle of H is read */
    double l00 = sqrt(H[0][0]);
    double l10 = H[1][0] / l00;
    double l20 = H[2][0] / l00;
    double l11 = sqrt(fma(-l10, l10, H[1][1]));
    double l21 = fma(-l20, l10, H[2][1]) / l11;
    double l22 = sqrt(fma(-l21, l21, fma(-l20, l20, H[2][2])));
    L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
    return (l00 > 0.0 && l11 > 0.0 && l22 > 0.0) ? 0 : 1;   /* NaN fails the comparisons */
}

static void chol3_solve(const double L[6], const double b[3], double x[3])
{
    double y0 = b[0] / L[0];
    double y1 = fma(-L[1], y0, b[1]) / L[2];
    double y2 = fma(-L[4], y1, fma(-L[3], y0, b[2])) / L[5];
    double x2 = y2 / L[5];
    double x1 = fma(-L[4], x2, y1) / L[2];
    double x0 = fma(-L[3], x2, fma(-L[1], x1, y0)) / L[0];
    x[0] = x0; x[1] = x1; x[2] = x2;
}

/* ------------------------------------------------------------------ a1: riccati_factor */
/* A..R point at THIS problem's (or the shared) model.  fac: N records of FS doubles. */
static int riccati_factor(int N, const double *A, const double *B, const double *c,
                          const double *Q, const double *R, double rho,
                          const int32_t *bt, double *fac)
{
    double P[6][6], Pk[6][6], T1[6][6], BtP[3][6], G1[3][6], H[3][3], L[6];
    const double rinv = 1.0 / rho;
    int bad = 0;
    for (int r = 0; r < 6; ++r)
        for (int i = 0; i < 6; ++i) {
            double w = (r == i && bt[3 * N + r / 3] != BLK_NONE) ? 1.0 : 0.0;
            if (Q) w = fma(Q[36 * N + r + 6 * i], rinv, w);
            P[r][i] = w;
        }
    for (int k = N - 1; k >= 0; --k) {
        double *f = fac + (size_t)FS * k;
        const double *Ak = A + 36 * k, *Bk = B + 18 * k;       /* col-major: A[i + 6 j] */
        double (*Am)[6] = (double (*)[6])(f + F_A);
        double (*Bm)[3] = (double (*)[3])(f + F_B);
        double (*Km)[6] = (double (*)[6])(f + F_K);
        double (*Acl)[6] = (double (*)[6])(f + F_ACL);
        double (*Hi)[3] = (double (*)[3])(f + F_HINV);
        double (*Em)[6] = (double (*)[6])(f + F_E);
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) Am[i][j] = Ak[i + 6 * j];
            for (int j = 0; j < 3; ++j) Bm[i][j] = Bk[i + 6 * j];
        }
        for (int j = 0; j < 3; ++j)
            for (int l = 0; l < 6; ++l) {
                double acc = Bm[0][j] * P[0][l];
                for (int i = 1; i < 6; ++i) acc = fma(Bm[i][j], P[i][l], acc);
                BtP[j][l] = acc;
            }
        const double wc = (bt[3 * k + 2] != BLK_NONE) ? 1.0 : 0.0;
        for (int j = 0; j < 3; ++j)
            for (int m = 0; m <= j; ++m) {
                double acc = (j == m) ? wc : 0.0;
                if (R) acc = fma(R[9 * k + j + 3 * m], rinv, acc);
                for (int l = 0; l < 6; ++l) acc = fma(BtP[j][l], Bm[l][m], acc);
                H[j][m] = acc; H[m][j] = acc;
            }
        bad |= chol3(H, L);
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 6; ++i) {
                double acc = BtP[j][0] * Am[0][i];
                for (int l = 1; l < 6; ++l) acc = fma(BtP[j][l], Am[l][i], acc);
                G1[j][i] = acc;
            }
        for (int i = 0; i < 6; ++i) {
            double b[3] = { G1[0][i], G1[1][i], G1[2][i] }, x[3];
            chol3_solve(L, b, x);
            Km[0][i] = -x[0]; Km[1][i] = -x[1]; Km[2][i] = -x[2];
            double b2[3] = { Bm[i][0], Bm[i][1], Bm[i][2] };
            chol3_solve(L, b2, x);
            Em[0][i] = x[0]; Em[1][i] = x[1]; Em[2][i] = x[2];
        }
        for (int m = 0; m < 3; ++m) {
            double b[3] = { m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0 }, x[3];
            chol3_solve(L, b, x);
            Hi[0][m] = x[0]; Hi[1][m] = x[1]; Hi[2][m] = x[2];
        }
        for (int l = 0; l < 6; ++l)
            for (int i = 0; i < 6; ++i) {
                double acc = Am[l][i];
                for (int j = 0; j < 3; ++j) acc = fma(Bm[l][j], Km[j][i], acc);
                Acl[l][i] = acc;
            }
        for (int r = 0; r < 6; ++r) {
            double cv = c ? c[6 * k + r] : 0.0;
            f[F_C + r] = cv;
        }
        for (int r = 0; r < 6; ++r) {
            double acc = P[r][0] * f[F_C + 0];
            for (int l = 1; l < 6; ++l) acc = fma(P[r][l], f[F_C + l], acc);
            f[F_CHAT + r] = acc;
        }
        f[FS - 1] = 0.0;
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 6; ++i) {
                double acc = P[r][0] * Acl[0][i];
                for (int l = 1; l < 6; ++l) acc = fma(P[r][l], Acl[l][i], acc);
                T1[r][i] = acc;
            }
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 6; ++i) {
                double acc = (r == i && bt[3 * k + r / 3] != BLK_NONE) ? 1.0 : 0.0;
                if (Q) acc = fma(Q[36 * k + r + 6 * i], rinv, acc);
                for (int l = 0; l < 6; ++l) acc = fma(Am[l][r], T1[l][i], acc);
                Pk[r][i] = acc;
            }
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 6; ++i) P[r][i] = 0.5 * (Pk[r][i] + Pk[i][r]);
    }
    return bad;
}

/* ------------------------------------------------------------------ a2: xupdate_riccati */
/* rt: right-hand side w*(z-u) - q/rho (n), d: scratch (3N), x: out (n). has_c: use c/chat. */
static void xupdate_riccati(int N, const double *fac, int has_c, const double *s0,
                            const double *rt, double *d, double *x)
{
    double g[6], pn[6];
    for (int i = 0; i < 6; ++i) g[i] = rt[9 * N + i];
    for (int k = N - 1; k >= 0; --k) {
        const double *f = fac + (size_t)FS * k;
        const double (*Km)[6] = (const double (*)[6])(f + F_K);
        const double (*Acl)[6] = (const double (*)[6])(f + F_ACL);
        const double (*Hi)[3] = (const double (*)[3])(f + F_HINV);
        const double (*Em)[6] = (const double (*)[6])(f + F_E);
        const double *rs = rt + 9 * k, *ra = rt + 9 * k + 6;
        if (has_c) for (int i = 0; i < 6; ++i) g[i] = g[i] - f[F_CHAT + i];
        for (int j = 0; j < 3; ++j) {
            double acc = Hi[j][0] * ra[0];
            acc = fma(Hi[j][1], ra[1], acc);
            acc = fma(Hi[j][2], ra[2], acc);
            for (int i = 0; i < 6; ++i) acc = fma(Em[j][i], g[i], acc);
            d[3 * k + j] = acc;
        }
        for (int i = 0; i < 6; ++i) {
            double acc = rs[i];
            for (int j = 0; j < 3; ++j) acc = fma(Km[j][i], ra[j], acc);
            for (int l = 0; l < 6; ++l) acc = fma(Acl[l][i], g[l], acc);
            pn[i] = acc;
        }
        for (int i = 0; i < 6; ++i) g[i] = pn[i];
    }
    double s[6], sn[6], a[3];
    for (int i = 0; i < 6; ++i) s[i] = s0[i];
    for (int k = 0; k < N; ++k) {
        const double *f = fac + (size_t)FS * k;
        const double (*Km)[6] = (const double (*)[6])(f + F_K);
        const double (*Am)[6] = (const double (*)[6])(f + F_A);
        const double (*Bm)[3] = (const double (*)[3])(f + F_B);
        for (int j = 0; j < 3; ++j) {
            double acc = d[3 * k + j];
            for (int i = 0; i < 6; ++i) acc = fma(Km[j][i], s[i], acc);
            a[j] = acc;
        }
        for (int i = 0; i < 6; ++i) x[9 * k + i] = s[i];
        for (int j = 0; j < 3; ++j) x[9 * k + 6 + j] = a[j];
        for (int i = 0; i < 6; ++i) {
            double acc = Am[i][0] * s[0];
            for (int l = 1; l < 6; ++l) acc = fma(Am[i][l], s[l], acc);
            for (int j = 0; j < 3; ++j) acc = fma(Bm[i][j], a[j], acc);
            if (has_c) acc = acc + f[F_C + i];
            sn[i] = acc;
        }
        for (int i = 0; i < 6; ++i) s[i] = sn[i];
    }
    for (int i = 0; i < 6; ++i) x[9 * N + i] = s[i];
}

/* ------------------------------------------------------------------ a1' / a2': dense shared factor */
/* M [n x n] row-major (M[i*ld + l]), S [n x 6], mc [n]; built from unit-vector Riccati solves
 * (column i of M is the x-update of rt = e_i with s_init = 0, c = 0). */
static void kkt_dense_factor(int N, const double *fac, int has_c, double *M, double *S, double *mc)
{
    const int n = 9 * N + 6;
    double *rt = (double *)calloc((size_t)n, sizeof(double));
    double *d = (double *)calloc((size_t)3 * N, sizeof(double));
    double *x = (double *)calloc((size_t)n, sizeof(double));
    double s0[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < n; ++i) {
        rt[i] = 1.0;
        xupdate_riccati(N, fac, 0, s0, rt, d, x);
        for (int r = 0; r < n; ++r) M[(size_t)r * n + i] = x[r];
        rt[i] = 0.0;
    }
    for (int j = 0; j < 6; ++j) {
        s0[j] = 1.0;
        xupdate_riccati(N, fac, 0, s0, rt, d, x);
        for (int r = 0; r < n; ++r) S[(size_t)r * 6 + j] = x[r];
        s0[j] = 0.0;
    }
    if (has_c) { xupdate_riccati(N, fac, 1, s0, rt, d, x); memcpy(mc, x, sizeof(double) * n); }
    else memset(mc, 0, sizeof(double) * n);
    free(rt); free(d); free(x);
}

static void xupdate_dense(int n, const double *M, const double *S, const double *mc,
                          const double *s0, const double *rt, double *x)
{
    for (int i = 0; i < n; ++i) {
        double acc = mc[i];
        for (int j = 0; j < 6; ++j) acc = fma(S[(size_t)i * 6 + j], s0[j], acc);
        const double *Mi = M + (size_t)i * n;
        for (int l = 0; l < n; ++l) acc = fma(Mi[l], rt[l], acc);
        x[i] = acc;
    }
}

/* ------------------------------------------------------------------ a3: prox of one 3-block */
static void prox_block(int type, const double *par, double rinv, const double v[3], double z[3])
{
    const double kap = par[PAR_LAM] * rinv;
    switch (type) {
    case BLK_L1:
    case BLK_L1_BOX:
        for (int e = 0; e < 3; ++e) {
            double t = v[e] > kap ? v[e] - kap : (v[e] < -kap ? v[e] + kap : 0.0);
            if (type == BLK_L1_BOX) {
                double lo = par[PAR_LO + e], hi = par[PAR_HI + e];
                t = t < lo ? lo : (t > hi ? hi : t);
            }
            z[e] = t;
        }
        break;
    case BLK_L2:
    case BLK_L2_BALL: {
        double sq = v[0] * v[0];
        sq = fma(v[1], v[1], sq);
        sq = fma(v[2], v[2], sq);
        double nrm = sqrt(sq);
        if (nrm > kap) {
            double mag = nrm - kap;
            if (type == BLK_L2_BALL && mag > par[PAR_RAD]) mag = par[PAR_RAD];
            double sc = mag / nrm;
            for (int e = 0; e < 3; ++e) z[e] = sc * v[e];
        } else {
            z[0] = z[1] = z[2] = 0.0;
        }
        break;
    }
    case BLK_BOX:
        for (int e = 0; e < 3; ++e) {
            double lo = par[PAR_LO + e], hi = par[PAR_HI + e];
            z[e] = v[e] < lo ? lo : (v[e] > hi ? hi : v[e]);
        }
        break;
    case BLK_BALL: {
        double w0 = v[0] - par[PAR_LO], w1 = v[1] - par[PAR_LO + 1], w2 = v[2] - par[PAR_LO + 2];
        double sq = w0 * w0;
        sq = fma(w1, w1, sq);
        sq = fma(w2, w2, sq);
        double nrm = sqrt(sq);
        if (nrm > par[PAR_RAD]) {
            double sc = par[PAR_RAD] / nrm;
            z[0] = fma(sc, w0, par[PAR_LO]);
            z[1] = fma(sc, w1, par[PAR_LO + 1]);
            z[2] = fma(sc, w2, par[PAR_LO + 2]);
        } else {
            z[0] = v[0]; z[1] = v[1]; z[2] = v[2];
        }
        break;
    }
    case BLK_POINT:
        z[0] = par[PAR_LO]; z[1] = par[PAR_LO + 1]; z[2] = par[PAR_LO + 2];
        break;
    default: /* BLK_FREE */
        z[0] = v[0]; z[1] = v[1]; z[2] = v[2];
        break;
    }
}

/* ------------------------------------------------------------------ a3+a4 over all blocks */
/* In-place on z,u (split entries only; BLK_NONE entries are left untouched).  norms[5] =
 * |x-z|^2, |z-z_old|^2, |x|^2, |z|^2, |u|^2 over split entries, sequential block order. */
static void prox_dual_residuals(int nb, const int32_t *bt, const double *par /* [8 x nb] */,
                                double rinv, double alpha, const double *x,
                                double *z, double *u, double norms[5])
{
    const double oma = 1.0 - alpha;
    double rr = 0.0, ss = 0.0, xx = 0.0, zz = 0.0, uu = 0.0;
    for (int b = 0; b < nb; ++b) {
        if (bt[b] == BLK_NONE) continue;
        double v[3], zn[3];
        for (int e = 0; e < 3; ++e) {
            int i = 3 * b + e;
            double xh = fma(alpha, x[i], oma * z[i]);
            v[e] = xh + u[i];
        }
        prox_block(bt[b], par + 8 * b, rinv, v, zn);
        for (int e = 0; e < 3; ++e) {
            int i = 3 * b + e;
            double un = v[e] - zn[e];
            double dr = x[i] - zn[e];
            double ds = zn[e] - z[i];
            rr = fma(dr, dr, rr);
            ss = fma(ds, ds, ss);
            xx = fma(x[i], x[i], xx);
            zz = fma(zn[e], zn[e], zz);
            uu = fma(un, un, uu);
            z[i] = zn[e];
            u[i] = un;
        }
    }
    norms[0] = rr; norms[1] = ss; norms[2] = xx; norms[3] = zz; norms[4] = uu;
}

/* ------------------------------------------------------------------ a5: adapt_rho */
#define RHO_MAX 1.0e6   /* rho is never pushed outside [RHO_MIN, RHO_MAX] (infeasible problems) */
#define RHO_MIN 1.0e-6
static int adapt_rho(double r_norm, double s_norm, double mu, double tau, double inv_tau,
                     double *rho, double *u_scale)
{
    if (r_norm > mu * s_norm) {
        if (*rho * tau > RHO_MAX) return 0;
        *rho = *rho * tau; *u_scale = inv_tau; return 1;
    }
    if (s_norm > mu * r_norm) {
        if (*rho * inv_tau < RHO_MIN) return 0;
        *rho = *rho * inv_tau; *u_scale = tau; return 1;
    }
    return 0;
}

static double now_seconds(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* ------------------------------------------------------------------ a6: the driver */
int ocp_solve(const ocp_problem *pb, const ocp_opts *op, ocp_result *res, int nthreads)
{
    const int N = pb->N;
    const int n = 9 * N + 6, nb = 3 * N + 2;
    const int64_t Bsz = pb->batch;
    const int has_c = pb->c != NULL, has_P = (pb->Q != NULL) || (pb->R != NULL);
    const int use_dense = (op->xupdate == 1);
    const int per_rho = op->adapt_rho || pb->rho0 != NULL;
    const int32_t *bt = pb->block_type;
    int nsplit = 0;
    for (int b = 0; b < nb; ++b) nsplit += (bt[b] != BLK_NONE) ? 3 : 0;
    if (use_dense && (pb->dyn_batched || (has_P && per_rho))) return -1;
    /* one factor for the whole batch iff the model is shared and (P == 0 or one shared rho) */
    const int shared_factor = !pb->dyn_batched && (!has_P || !per_rho);
    const double sqrtn_abs = sqrt((double)nsplit) * op->abstol;
    const double inv_tau = 1.0 / op->adapt_tau;
    double *fac_sh = NULL, *M = NULL, *S = NULL, *mc = NULL;
    int rc = 0;
    if (shared_factor) {
        fac_sh = (double *)malloc(sizeof(double) * FS * (size_t)N);
        if (riccati_factor(N, pb->A, pb->B, pb->c, pb->Q, pb->R, op->rho, bt, fac_sh)) rc = 1;
        if (use_dense) {
            M = (double *)malloc(sizeof(double) * (size_t)n * n);
            S = (double *)malloc(sizeof(double) * (size_t)n * 6);
            mc = (double *)malloc(sizeof(double) * (size_t)n);
            kkt_dense_factor(N, fac_sh, has_c, M, S, mc);
        }
    }
    int64_t n_conv = 0, sum_it = 0, max_it = 0, n_refac = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const double t0 = now_seconds();
#pragma omp parallel reduction(+ : n_conv, sum_it, n_refac) reduction(max : max_it)
    {
        double *x = (double *)malloc(sizeof(double) * n);
        double *z = (double *)malloc(sizeof(double) * n);
        double *u = (double *)malloc(sizeof(double) * n);
        double *rt = (double *)malloc(sizeof(double) * n);
        double *d = (double *)malloc(sizeof(double) * 3 * (size_t)N);
        double *fac_own = shared_factor ? NULL : (double *)malloc(sizeof(double) * FS * (size_t)N);
#pragma omp for schedule(dynamic, 4)
        for (int64_t p = 0; p < Bsz; ++p) {
            const size_t pd = pb->dyn_batched ? (size_t)p : 0;
            const double *Ap = pb->A + 36 * (size_t)N * pd, *Bp = pb->B + 18 * (size_t)N * pd;
            const double *cp = has_c ? pb->c + 6 * (size_t)N * pd : NULL;
            const double *Qp = pb->Q ? pb->Q + 36 * (size_t)(N + 1) * pd : NULL;
            const double *Rp = pb->R ? pb->R + 9 * (size_t)N * pd : NULL;
            const double *qp = pb->q ? pb->q + (pb->q_batched ? (size_t)n * p : 0) : NULL;
            const double *par = pb->block_par + (pb->par_batched ? (size_t)8 * nb * p : 0);
            const double *s0 = pb->s0 + 6 * (size_t)p;
            double rho = pb->rho0 ? pb->rho0[p] : op->rho;
            const double *fac = fac_sh;
            int st = ST_MAX_ITER, it = 0;
            if (!shared_factor) {
                if (riccati_factor(N, Ap, Bp, cp, Qp, Rp, rho, bt, fac_own)) st = ST_NAN;
                fac = fac_own;
            }
            for (int i = 0; i < n; ++i) {
                int split = bt[i / 3] != BLK_NONE;
                z[i] = (split && pb->z0) ? pb->z0[(size_t)n * p + i] : 0.0;
                u[i] = (split && pb->u0) ? pb->u0[(size_t)n * p + i] : 0.0;
                x[i] = 0.0;
            }
            double r_norm = 0.0, s_norm = 0.0, eps_pri = 0.0, eps_dual = 0.0;
            if (st != ST_NAN)
            for (it = 1; it <= op->max_iter; ++it) {
                const double rinv = 1.0 / rho;
                for (int i = 0; i < n; ++i) {
                    double t;
                    if (bt[i / 3] != BLK_NONE) {
                        t = z[i] - u[i];
                        if (qp) t = fma(-qp[i], rinv, t);
                    } else {
                        t = qp ? -(qp[i] * rinv) : 0.0;
                    }
                    rt[i] = t;
                }
                if (use_dense) xupdate_dense(n, M, S, mc, s0, rt, x);
                else xupdate_riccati(N, fac, has_c, s0, rt, d, x);
                double nr[5];
                prox_dual_residuals(nb, bt, par, rinv, op->alpha, x, z, u, nr);
                r_norm = sqrt(nr[0]);
                s_norm = rho * sqrt(nr[1]);
                double nx = sqrt(nr[2]), nz = sqrt(nr[3]);
                eps_pri = fma(op->reltol, nx > nz ? nx : nz, sqrtn_abs);
                eps_dual = fma(op->reltol, rho * sqrt(nr[4]), sqrtn_abs);
                if (res->hist_r) {
                    size_t h = (size_t)op->max_iter * p + (it - 1);
                    res->hist_r[h] = r_norm; res->hist_s[h] = s_norm;
                    res->hist_eps_pri[h] = eps_pri; res->hist_eps_dual[h] = eps_dual;
                    res->hist_rho[h] = rho;
                }
                if (!(isfinite(r_norm) && isfinite(s_norm))) { st = ST_NAN; break; }
                if (r_norm < eps_pri && s_norm < eps_dual) { st = ST_CONVERGED; break; }
                if (op->adapt_rho && (it % op->adapt_every) == 0 && it < op->max_iter &&
                    (op->adapt_until <= 0 || it <= op->adapt_until)) {
                    double usc;
                    if (adapt_rho(r_norm, s_norm, op->adapt_mu, op->adapt_tau, inv_tau, &rho, &usc)) {
                        for (int i = 0; i < n; ++i) u[i] = u[i] * usc;
                        if (has_P) {
                            if (riccati_factor(N, Ap, Bp, cp, Qp, Rp, rho, bt, fac_own)) { st = ST_NAN; break; }
                            ++n_refac;
                        }
                    }
                }
            }
            if (it > op->max_iter) it = op->max_iter;
            if (res->hist_r)
                for (int k = it; k < op->max_iter; ++k) {
                    size_t h = (size_t)op->max_iter * p + k;
                    res->hist_r[h] = NAN; res->hist_s[h] = NAN; res->hist_eps_pri[h] = NAN;
                    res->hist_eps_dual[h] = NAN; res->hist_rho[h] = NAN;
                }
            for (int i = 0; i < n; ++i) {
                int split = bt[i / 3] != BLK_NONE;
                if (res->x) res->x[(size_t)n * p + i] = x[i];
                if (res->z) res->z[(size_t)n * p + i] = split ? z[i] : x[i];
                if (res->u) res->u[(size_t)n * p + i] = split ? u[i] : 0.0;
            }
            res->iters[p] = it; res->status[p] = st;
            if (res->r_norm) res->r_norm[p] = r_norm;
            if (res->s_norm) res->s_norm[p] = s_norm;
            if (res->eps_pri) res->eps_pri[p] = eps_pri;
            if (res->eps_dual) res->eps_dual[p] = eps_dual;
            if (res->rho) res->rho[p] = rho;
            n_conv += (st == ST_CONVERGED);
            sum_it += it;
            if (it > max_it) max_it = it;
        }
        free(x); free(z); free(u); free(rt); free(d); free(fac_own);
    }
    res->seconds = now_seconds() - t0;
    res->stats[0] = n_conv; res->stats[1] = sum_it; res->stats[2] = max_it; res->stats[3] = n_refac;
    free(fac_sh); free(M); free(S); free(mc);
    return rc;
}

/* ------------------------------------------------------------------ unit entry points (tests) */
int ocp_riccati_factor(int N, const double *A, const double *B, const double *c, const double *Q,
                       const double *R, double rho, const int32_t *bt, double *fac)
{ return riccati_factor(N, A, B, c, Q, R, rho, bt, fac); }

void ocp_xupdate_riccati(int N, const double *fac, int has_c, const double *s0, const double *rt,
                         double *x)
{
    double *d = (double *)malloc(sizeof(double) * 3 * (size_t)N);
    xupdate_riccati(N, fac, has_c, s0, rt, d, x);
    free(d);
}

void ocp_kkt_dense_factor(int N, const double *fac, int has_c, double *M, double *S, double *mc)
{ kkt_dense_factor(N, fac, has_c, M, S, mc); }

void ocp_xupdate_dense(int n, const double *M, const double *S, const double *mc,
                       const double *s0, const double *rt, double *x)
{ xupdate_dense(n, M, S, mc, s0, rt, x); }

void ocp_prox_dual_residuals(int nb, const int32_t *bt, const double *par, double rinv, double alpha,
                             const double *x, double *z, double *u, double *norms)
{ prox_dual_residuals(nb, bt, par, rinv, alpha, x, z, u, norms); }

int ocp_factor_stride(void) { return FS; }
int ocp_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
