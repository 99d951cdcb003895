// admm_mex.cpp -- the thin MEX gateway  [x, z, u, hist] = admm_mex(prob, opts [, scp])
// (layer L2 of SURVEY.md 1.2; BASELINE.json north_star: "MATLAB host code calls CUDA through a thin
// C-ABI MEX layer, with no gpuArray").  It validates mxArray fields, extracts raw pointers
// (zero-copy on the host side: MATLAB's column-major arrays ARE the C-ABI layout), allocates the
// outputs and calls admmb_solve -- or, with a third argument, admmb_scp_solve (the stage matrices are
// then re-linearised on the device each pass, prob carries N instead of A, B).  No math happens here.
//
// Build (on a machine that has MATLAB):  mex -R2018a admm_mex.cpp -I../../include -L../lib -ladmm_b200
// In this repository it is only syntax-checked against mex/stub/mex.h (no MATLAB in the image).
#include <cstring>
#include <string>

#include "mex.h"
#include "admm_b200.h"

namespace {

admmb_handle g_handle = nullptr;
int g_gpus = 0;

void at_exit()
{
    if (g_handle) admmb_destroy(g_handle);
    g_handle = nullptr;
}

const char *code_name(int rc)
{
    switch (rc) {
    case ADMMB_E_BADARG: return "admm:ADMMB_E_BADARG";
    case ADMMB_E_CUDA: return "admm:ADMMB_E_CUDA";
    case ADMMB_E_NCCL: return "admm:ADMMB_E_NCCL";
    case ADMMB_E_NOMEM: return "admm:ADMMB_E_NOMEM";
    case ADMMB_E_NODEVICE: return "admm:ADMMB_E_NODEVICE";
    case ADMMB_E_STATE: return "admm:ADMMB_E_STATE";
    default: return "admm:unknown";
    }
}

const mxArray *field(const mxArray *s, const char *name, bool required)
{
    const mxArray *f = mxGetField(s, 0, name);
    if ((!f || mxIsEmpty(f)) && required) mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "prob.%s is required", name);
    return (f && !mxIsEmpty(f)) ? f : nullptr;
}

const double *dptr(const mxArray *a, const char *name)
{
    if (!a) return nullptr;
    if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "%s must be real double", name);
    return mxGetDoubles(a);
}

size_t dim(const mxArray *a, size_t i)
{
    return i < mxGetNumberOfDimensions(a) ? mxGetDimensions(a)[i] : 1;
}

double opt_scalar(const mxArray *o, const char *name, double dflt)
{
    const mxArray *f = o ? mxGetField(o, 0, name) : nullptr;
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

int opt_enum(const mxArray *o, const char *name, const char *const *names, int count, int dflt)
{
    const mxArray *f = o ? mxGetField(o, 0, name) : nullptr;
    if (!f || mxIsEmpty(f)) return dflt;
    if (!mxIsChar(f)) return (int)mxGetScalar(f);
    char buf[32];
    mxGetString(f, buf, sizeof(buf));
    for (int i = 0; i < count; ++i)
        if (std::strcmp(buf, names[i]) == 0) return i;
    mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "opts.%s = '%s' is not recognised", name, buf);
    return dflt;
}

mxArray *new_double(size_t rows, size_t cols, double **p)
{
    size_t dims[2] = {rows, cols};
    mxArray *a = mxCreateUninitNumericArray(2, dims, mxDOUBLE_CLASS, mxREAL);
    *p = mxGetDoubles(a);
    return a;
}

}  // namespace

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs < 1 || nrhs > 3 || !mxIsStruct(prhs[0]) || (nrhs >= 2 && !mxIsStruct(prhs[1])) || (nrhs == 3 && !mxIsStruct(prhs[2])))
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "usage: [x,z,u,hist] = admm_mex(prob, opts [, scp])");
    if (nlhs > 4) mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "at most four outputs");
    const mxArray *P = prhs[0], *O = nrhs >= 2 ? prhs[1] : nullptr, *S = nrhs == 3 ? prhs[2] : nullptr;

    // ---- problem: raw pointers into MATLAB's own (read-only) buffers
    admmb_problem pb;
    std::memset(&pb, 0, sizeof(pb));
    const mxArray *A = field(P, "A", !S), *B = field(P, "B", !S), *s0 = field(P, "s0", true);
    if (S) A = B = nullptr;                               // SCP: the model is produced on the device
    const mxArray *bt = field(P, "block_type", true), *bp = field(P, "block_par", true);
    const mxArray *c = field(P, "c", false), *Q = field(P, "Q", false), *R = field(P, "R", false);
    const mxArray *q = field(P, "q", false), *z0 = field(P, "z0", false), *u0 = field(P, "u0", false);
    const mxArray *rho0 = field(P, "rho0", false);
    if (dim(s0, 0) != 6 || (!S && (dim(A, 0) != 6 || dim(A, 1) != 6 || dim(B, 0) != 6 || dim(B, 1) != 3)))
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "A must be 6x6xNxBd, B 6x3xNxBd, s0 6xBsz");
    const size_t Bsz = dim(s0, 1);
    const size_t N = S ? (size_t)mxGetScalar(field(P, "N", true)) : dim(A, 2), Bd = S ? Bsz : dim(A, 3);
    const size_t n = 9 * N + 6, nb = 3 * N + 2;
    if (!S && (dim(B, 2) != N || dim(B, 3) != Bd || (Bd != 1 && Bd != Bsz)))
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "A and B must agree in N and be shared (Bd=1) or per problem (Bd=Bsz)");
    if (!mxIsInt32(bt) || mxGetNumberOfElements(bt) != nb)
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "block_type must be int32 with 3N+2 entries");
    if (dim(bp, 0) != 8 || dim(bp, 1) != nb || (dim(bp, 2) != 1 && dim(bp, 2) != Bsz))
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "block_par must be 8 x (3N+2) x (1 or Bsz)");
    auto check = [&](const mxArray *a, size_t want, const char *name) {
        if (a && mxGetNumberOfElements(a) != want) mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "prob.%s has the wrong size", name);
    };
    check(c, 6 * N * Bd, "c");
    check(Q, 36 * (N + 1) * Bd, "Q");
    check(R, 9 * N * Bd, "R");
    check(z0, n * Bsz, "z0");
    check(u0, n * Bsz, "u0");
    check(rho0, Bsz, "rho0");
    if (q && mxGetNumberOfElements(q) != n && mxGetNumberOfElements(q) != n * Bsz)
        mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "prob.q must be n x 1 or n x Bsz");
    pb.N = (int32_t)N;
    pb.batch = (int64_t)Bsz;
    pb.A = dptr(A, "A"); pb.B = dptr(B, "B"); pb.c = S ? nullptr : dptr(c, "c"); pb.Q = dptr(Q, "Q"); pb.R = dptr(R, "R");
    pb.dyn_batched = Bd > 1;
    pb.q = dptr(q, "q");
    pb.q_batched = q && mxGetNumberOfElements(q) == n * Bsz && Bsz > 1;
    pb.s0 = dptr(s0, "s0");
    pb.block_type = mxGetInt32s(bt);
    pb.block_par = dptr(bp, "block_par");
    pb.par_batched = dim(bp, 2) > 1;
    pb.z0 = dptr(z0, "z0"); pb.u0 = dptr(u0, "u0"); pb.rho0 = dptr(rho0, "rho0");

    // ---- options
    static const char *const xu[] = {"auto", "dense", "riccati"};
    static const char *const pr[] = {"fp64", "tf32"};
    admmb_opts op;
    std::memset(&op, 0, sizeof(op));
    op.rho = opt_scalar(O, "rho", 1.0);
    op.alpha = opt_scalar(O, "alpha", 1.0);
    op.abstol = opt_scalar(O, "abstol", 1e-6);
    op.reltol = opt_scalar(O, "reltol", 1e-6);
    op.max_iter = (int32_t)opt_scalar(O, "max_iter", 1000);
    op.adapt_rho = (int32_t)opt_scalar(O, "adapt_rho", 0);
    op.adapt_mu = opt_scalar(O, "adapt_mu", 10.0);
    op.adapt_tau = opt_scalar(O, "adapt_tau", 2.0);
    op.adapt_every = (int32_t)opt_scalar(O, "adapt_every", 25);
    op.adapt_until = (int32_t)opt_scalar(O, "adapt_until", 0);
    op.xupdate = opt_enum(O, "xupdate", xu, 3, ADMMB_XUPDATE_AUTO);
    op.precision = opt_enum(O, "precision", pr, 2, ADMMB_PREC_FP64);
    op.history = (int32_t)opt_scalar(O, "history", 0);
    op.chunk = (int32_t)opt_scalar(O, "chunk", 0);
    op.kernel = (int32_t)opt_scalar(O, "kernel", 0);
    op.tf32_switch = (int32_t)opt_scalar(O, "tf32_switch", 0);
    op.tf32_refresh = (int32_t)opt_scalar(O, "tf32_refresh", 0);
    const int gpus = (int)opt_scalar(O, "gpus", 0);

    // ---- SCP parameters (third argument): model, stage length, orbit radius, passes, trajectory tolerances
    admmb_scp sc;
    std::memset(&sc, 0, sizeof(sc));
    if (S) {
        static const char *const models[] = {"", "nl_circular", "nl_elliptic"};
        sc.model = opt_enum(S, "model", models, 3, ADMMB_SCP_NL_CIRCULAR);
        const mxArray *se = mxGetField(S, 0, "e"), *sth = mxGetField(S, 0, "theta0");
        if (sc.model == ADMMB_SCP_NL_ELLIPTIC) {
            if (!se || !sth || mxGetNumberOfElements(se) != Bsz || mxGetNumberOfElements(sth) != Bsz)
                mexErrMsgIdAndTxt("admm:ADMMB_E_BADARG", "scp.e and scp.theta0 must have one entry per problem");
            sc.e = dptr(se, "scp.e");
            sc.theta0 = dptr(sth, "scp.theta0");
        }
        sc.substeps = (int32_t)opt_scalar(S, "substeps", 0);
        sc.T = opt_scalar(S, "T", 0.0);
        sc.nmm = opt_scalar(S, "nmm", 0.0);
        sc.R0 = opt_scalar(S, "R0", 0.0);
        sc.max_pass = (int32_t)opt_scalar(S, "max_pass", 10);
        sc.tol_abs = opt_scalar(S, "tol_abs", 0.0);
        sc.tol_rel = opt_scalar(S, "tol_rel", 1e-6);
        static const char *const ctrls[] = {"zoh", "impulsive"};
        sc.control = opt_enum(S, "control", ctrls, 2, ADMMB_SCP_CTRL_ZOH);
    }

    // ---- handle: created once, kept across calls (CUDA context, device buffers, worker threads)
    if (g_handle && gpus != g_gpus) at_exit();
    if (!g_handle) {
        int rc = admmb_create(&g_handle, nullptr, gpus);
        if (rc != ADMMB_OK) mexErrMsgIdAndTxt(code_name(rc), "%s", admmb_last_error(nullptr));
        g_gpus = gpus;
        mexAtExit(at_exit);
        mexLock();
    }

    // ---- outputs: allocated by MATLAB, filled by the library through their data pointers
    admmb_result res;
    std::memset(&res, 0, sizeof(res));
    plhs[0] = new_double(n, Bsz, &res.x);
    mxArray *mz = new_double(n, Bsz, &res.z), *mu = new_double(n, Bsz, &res.u);
    static const char *names[] = {"iters", "status", "r_norm", "s_norm", "eps_pri", "eps_dual", "rho", "hist_r_norm",
                                  "hist_s_norm", "hist_eps_pri", "hist_eps_dual", "hist_rho", "stats", "device_ms",
                                  "scp_passes", "scp_status", "scp_step", "scp_iters_total", "scp_hist_step", "scp_stats"};
    mxArray *H = mxCreateStructMatrix(1, 1, 20, names);
    mxArray *it = mxCreateNumericMatrix(Bsz, 1, mxINT32_CLASS, mxREAL), *st = mxCreateNumericMatrix(Bsz, 1, mxINT32_CLASS, mxREAL);
    res.iters = mxGetInt32s(it);
    res.status = mxGetInt32s(st);
    mxSetField(H, 0, "iters", it);
    mxSetField(H, 0, "status", st);
    mxSetField(H, 0, "r_norm", new_double(Bsz, 1, &res.r_norm));
    mxSetField(H, 0, "s_norm", new_double(Bsz, 1, &res.s_norm));
    mxSetField(H, 0, "eps_pri", new_double(Bsz, 1, &res.eps_pri));
    mxSetField(H, 0, "eps_dual", new_double(Bsz, 1, &res.eps_dual));
    mxSetField(H, 0, "rho", new_double(Bsz, 1, &res.rho));
    if (op.history) {
        mxSetField(H, 0, "hist_r_norm", new_double((size_t)op.max_iter, Bsz, &res.hist_r));
        mxSetField(H, 0, "hist_s_norm", new_double((size_t)op.max_iter, Bsz, &res.hist_s));
        mxSetField(H, 0, "hist_eps_pri", new_double((size_t)op.max_iter, Bsz, &res.hist_eps_pri));
        mxSetField(H, 0, "hist_eps_dual", new_double((size_t)op.max_iter, Bsz, &res.hist_eps_dual));
        mxSetField(H, 0, "hist_rho", new_double((size_t)op.max_iter, Bsz, &res.hist_rho));
    }

    int rc;
    if (S) {
        admmb_scp_result so;
        std::memset(&so, 0, sizeof(so));
        mxArray *ps = mxCreateNumericMatrix(Bsz, 1, mxINT32_CLASS, mxREAL), *ss = mxCreateNumericMatrix(Bsz, 1, mxINT32_CLASS, mxREAL);
        mxArray *ti = mxCreateNumericMatrix(Bsz, 1, mxINT64_CLASS, mxREAL);
        so.passes = mxGetInt32s(ps);
        so.scp_status = mxGetInt32s(ss);
        so.iters_total = mxGetInt64s(ti);
        mxSetField(H, 0, "scp_passes", ps);
        mxSetField(H, 0, "scp_status", ss);
        mxSetField(H, 0, "scp_iters_total", ti);
        mxSetField(H, 0, "scp_step", new_double(Bsz, 1, &so.step));
        mxSetField(H, 0, "scp_hist_step", new_double((size_t)(sc.max_pass > 0 ? sc.max_pass : 1), Bsz, &so.hist_step));
        rc = admmb_scp_solve(g_handle, &pb, &sc, &op, &res, &so);
        if (rc == ADMMB_OK) {
            mxArray *sst = mxCreateNumericMatrix(4, 1, mxINT64_CLASS, mxREAL);
            std::memcpy(mxGetInt64s(sst), so.stats, sizeof(so.stats));
            mxSetField(H, 0, "scp_stats", sst);
        }
    } else {
        rc = admmb_solve(g_handle, &pb, &op, &res);           // blocking; worker threads never touch mx*/mex*
    }
    if (rc != ADMMB_OK) mexErrMsgIdAndTxt(code_name(rc), "%s", admmb_last_error(g_handle));

    mxArray *stats = mxCreateNumericMatrix(4, 1, mxINT64_CLASS, mxREAL);
    std::memcpy(mxGetInt64s(stats), res.stats, sizeof(res.stats));
    mxSetField(H, 0, "stats", stats);
    double *ms;
    mxSetField(H, 0, "device_ms", new_double(1, 1, &ms));
    *ms = res.device_ms;
    if (nlhs > 1) plhs[1] = mz;
    if (nlhs > 2) plhs[2] = mu;
    if (nlhs > 3) plhs[3] = H;
}
