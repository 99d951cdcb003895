// dense_impl.cuh -- host side of the dense shared-factor path (included by admm_b200.cu).
#pragma once
#include <chrono>

namespace {

// Row a1': the dense factor is built ON DEVICE from the Riccati factor: column i of M is the
// x-update of the unit right-hand side e_i with s_init = 0 and c = 0, i.e. one batch of n+6(+1)
// "problems" pushed through the row-a2 kernel.
void dense_build_factor(Shard &s, const double *fac_dev, bool has_c, double *M, double *S, double *mc)
{
    const int n = s.n, N = s.N;
    const int64_t cols = n + 6 + 1;
    const size_t ldc = round_up((size_t)cols, 32);
    // persistent scratch: cudaMalloc / cudaFree per run cost erratic host time (cudaFree synchronises the device)
    DevBuf<double> &rt = s.dense.bf_rt, &s0 = s.dense.bf_s0, &d = s.dense.bf_d, &x = s.dense.bf_x;
    rt.alloc((size_t)n * ldc);
    s0.alloc(6 * ldc);
    d.alloc((size_t)3 * N * ldc);
    x.alloc((size_t)n * ldc);
    CK(cudaMemsetAsync(rt.p, 0, sizeof(double) * n * ldc, s.stream));
    CK(cudaMemsetAsync(s0.p, 0, sizeof(double) * 6 * ldc, s.stream));
    k_dense_unit_rhs<<<(unsigned)((cols + 127) / 128), 128, 0, s.stream>>>(n, ldc, rt.p, s0.p);
    ++s.launches;
    // columns 0..n+5: homogeneous dynamics (c = 0); column n+6: rt = 0, s0 = 0, with c
    k_xupdate_riccati<false><<<(unsigned)((n + 6 + 127) / 128), 128, 0, s.stream>>>(N, n + 6, ldc, fac_dev, s0.p, rt.p, d.p, x.p);
    ++s.launches;
    if (has_c) {
        k_xupdate_riccati<true><<<1, 32, 0, s.stream>>>(N, 1, ldc, fac_dev, s0.p + (n + 6), rt.p + (n + 6), d.p + (n + 6), x.p + (n + 6));
        ++s.launches;
    }
    CK(cudaGetLastError());
    // x is [n rows][ldc cols]: row r, column i = M[r][i]  -> already row-major with leading dim ldc
    k_dense_split_factor<<<(unsigned)((n + 127) / 128), 128, 0, s.stream>>>(n, ldc, x.p, has_c ? 1 : 0, M, S, mc);
    ++s.launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));
}

void dense_tf32_xupdate(Shard &s)
{
    s.dense.tf32.gemm(s.stream);
    ++s.launches;
}

// unit entry point: X = M RT + S s0 + mc through the tensor-core kernel, FP64 in / FP64 out
int dense_tf32_unit(Shard &s, int n, int64_t batch, size_t ld, const double *M, const double *S, const double *mc,
                    const double *s0, const double *rt, double *x, int split)
{
    Tf32Plan plan;
    plan.prepare(n, batch, ld, split, M, S, mc, s0, s.stream);
    dim3 g((unsigned)((ld + 127) / 128), (unsigned)n);
    k_tf32_split_rows<<<g, 128, 0, s.stream>>>(n, ld, rt, plan.Bhi.p, split == 3 ? plan.Blo.p : nullptr);
    plan.gemm(s.stream);
    k_tf32_to_double<<<g, 128, 0, s.stream>>>(n, ld, plan.X.p, x);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));
    return ADMMB_OK;
}

void dense_prepare(Shard &s, const admmb_opts *)
{
    DenseState &D = s.dense;
    D.M.alloc((size_t)s.n * s.n);
    D.S.alloc((size_t)s.n * 6);
    D.mc.alloc(s.n);
    if (s.use_dense) D.x.alloc((size_t)s.n * s.ld);
    D.running.alloc(1);
    D.ready = false;
}

// The condensed incremental TF32 loop (see dense_tf32.cuh) with physical compaction: the still-running problems
// travel in a dense, 32-aligned working set (the same ColArray machinery as the Riccati path: z, u, x_R, s0, rho,
// iters, status, fin, par and the increment buffers); a finished problem's full x is computed when it retires.
void condensed_loop(Shard &s, const admmb_opts *op, int split, int it0, bool tail, double t_setup);

// the split rows R and the list of split blocks
static void condensed_rows(Shard &s, std::vector<int> &R, std::vector<int> &sb)
{
    for (int b = 0; b < s.nb; ++b)
        if ((s.h_bdesc[b] & 0xff) != BLK_NONE) { sb.push_back(b); for (int e = 0; e < 3; ++e) R.push_back(3 * b + e); }
}

void dense_run_condensed(Shard &s, const admmb_opts *op, int split)
{
    DenseState &D = s.dense;
    Tf32Condensed &C = D.cond;
    const int n = s.n, nb = s.nb, N = s.N;
    const auto t_setup0 = std::chrono::steady_clock::now();
    const unsigned gb = (unsigned)((s.batch + 127) / 128);
    std::vector<int> R, sb;
    condensed_rows(s, R, sb);
    D.sblk.alloc(sb.size());
    CK(cudaMemcpyAsync(D.sblk.p, sb.data(), sizeof(int) * sb.size(), cudaMemcpyHostToDevice, s.stream));
    C.prepare(n, R, s.ld, split, D.M.p, s.stream);
    // x^1 exactly: one FP64 Riccati x-update of the initial right-hand side, then x_R = its split rows
    D.rt.alloc((size_t)n * s.ld);
    D.dscr.alloc((size_t)3 * N * s.ld);
    k_dense_rt_init<<<gb, 128, 0, s.stream>>>(nb, s.batch, s.ld, s.bdesc.p, s.z.p, s.u.p, s.rho.p, nullptr, 0, D.rt.p);
    if (s.has_c) k_xupdate_riccati<true><<<gb, 128, 0, s.stream>>>(N, s.batch, s.ld, s.fac.p, s.s0.p, D.rt.p, D.dscr.p, D.x.p);
    else k_xupdate_riccati<false><<<gb, 128, 0, s.stream>>>(N, s.batch, s.ld, s.fac.p, s.s0.p, D.rt.p, D.dscr.p, D.x.p);
    {
        dim3 g(gb, (unsigned)C.nr);
        k_tf32_gather_rows<<<g, 128, 0, s.stream>>>(C.nr, C.rows.p, s.batch, s.ld, D.x.p, C.Xacc.p);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));   // R, sb are host temporaries of the async copies above
    s.launches += 4;
    const double t_setup = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup0).count();

    for (int c = 0; c < Shard::C_COUNT; ++c) s.set_col(c, nullptr, 8, 0, false);
    s.set_col(Shard::C_Z, s.z.p, 8, n, true);
    s.set_col(Shard::C_U, s.u.p, 8, n, true);
    s.set_col(Shard::C_XACC, C.Xacc.p, 8, C.nr, false);
    s.set_col(Shard::C_S0, s.s0.p, 8, 6, false);
    s.set_col(Shard::C_RHO, s.rho.p, 8, 1, false);
    s.set_col(Shard::C_ITERS, s.iters.p, 4, 1, true);
    s.set_col(Shard::C_STATUS, s.status.p, 4, 1, true);
    s.set_col(Shard::C_FIN, s.fin.p, 8, 4, true);
    s.set_col(Shard::C_PAR, s.par_batched ? s.par.p : nullptr, 8, 8 * nb, false);
    s.set_col(Shard::C_BH, C.Bhi.p, 4, C.kpad, false);
    s.set_col(Shard::C_BL, split == 3 ? C.Blo.p : nullptr, 4, C.kpad, false);
    s.cur_set = -1;
    s.width = s.batch;
    s.ld_cur = s.ld;
    condensed_loop(s, op, split, 0, false, t_setup);
}

// precision = tf32, xupdate = auto: the Riccati loop hands over its working set (compact z / u rows, d, s0, rho,
// iters, status, fin, par already travelling) once it is narrow.  x_R starts from an exact FP64 x-update of the
// current (z, u); a finished problem leaves its d (backward-sweep result of its last right-hand side) in the
// working set, from which the Riccati path's k_output recomputes x at download exactly as for the other problems.
void dense_tail(Shard &s, const admmb_opts *op, int it0)
{
    DenseState &D = s.dense;
    Tf32Condensed &C = D.cond;
    const auto t_setup0 = std::chrono::steady_clock::now();
    const int split = getenv("ADMMB_TF32_SINGLE") ? 1 : 3;
    std::vector<int> R, sb;
    condensed_rows(s, R, sb);
    dense_build_factor(s, s.fac.p, s.has_c, D.M.p, D.S.p, D.mc.p);
    D.sblk.alloc(sb.size());
    CK(cudaMemcpyAsync(D.sblk.p, sb.data(), sizeof(int) * sb.size(), cudaMemcpyHostToDevice, s.stream));
    C.prepare(s.n, R, s.ld_cur, split, D.M.p, s.stream);      // buffers at the pitch of the current working set
    D.dscr.alloc((size_t)3 * s.N * s.ld_cur);
    CK(cudaStreamSynchronize(s.stream));
    s.set_col(Shard::C_XACC, C.Xacc.p, 8, C.nr, false);
    s.set_col(Shard::C_BH, C.Bhi.p, 4, C.kpad, false);
    s.set_col(Shard::C_BL, split == 3 ? C.Blo.p : nullptr, 4, C.kpad, false);
    s.cols[Shard::C_XACC].cur_override = C.Xacc.p;
    s.cols[Shard::C_BH].cur_override = C.Bhi.p;
    if (split == 3) s.cols[Shard::C_BL].cur_override = C.Blo.p;
    const double t_setup = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup0).count();
    condensed_loop(s, op, split, it0, true, t_setup);
}

// tail = false: z / u are full-row arrays, x_R is initialised by the caller, finished problems get their full x in D.x
// tail = true : z / u use the Riccati path's compact rows, x_R is initialised here from (z, u), finished problems
//               leave d in the working set
void condensed_loop(Shard &s, const admmb_opts *op, int split, int it0, bool tail, double t_setup)
{
    DenseState &D = s.dense;
    Tf32Condensed &C = D.cond;
    const int nb = s.nb, N = s.N;
    (void)nb;
    const bool no_repack = getenv("ADMMB_NO_REPACK") != nullptr;

    DenseStep ds{};
    ds.max_iter = op->max_iter;
    ds.reltol = op->reltol;
    ds.sqrtn_abs = sqrt((double)s.nsplit) * op->abstol;
    ds.running = D.running.p;
    int chunk = op->chunk > 0 ? op->chunk : 25;
    const int chunk_max = op->chunk > 0 ? op->chunk : 200;
    constexpr int GRAPH_ITERS = 10;
    const bool use_graphs = getenv("ADMMB_NO_GRAPH") == nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool gstale = false;
    struct GraphGuard { cudaGraphExec_t &g; ~GraphGuard() { if (g) cudaGraphExecDestroy(g); } } guard{gexec};
    D.itbase.alloc(1);
    const size_t dscr_ld = D.dscr.n / (size_t)(3 * N);      // pitch of the Riccati scratch
    // ADMMB_TRACE: host time per section of the loop (the GPU idles while the host works between chunks)
    const bool trace = getenv("ADMMB_TRACE") != nullptr;
    double t_sec[6] = {0, 0, 0, 0, 0, 0};   // launch, graph build, split+sync, final x, repack, bind
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    const int refresh = op->tf32_refresh > 0 ? op->tf32_refresh : (op->tf32_refresh < 0 ? 0 : 1000);      // opts.tf32_refresh
    const int zu_compact = tail ? 1 : 0;
    // the tail starts with a "refresh": x_R = exact x-update of the (z, u) the Riccati loop left, increment = 0
    // refresh schedule: early and often while the steps are large (that is where the increments' rounding accumulates:
    // it is relative to the step), then every `refresh` iterations
    int next_refresh = refresh > 0 ? it0 + 64 : 0x7fffffff;
    bool force_refresh = tail;
    int it = it0;
    while (s.width > 0 && it < op->max_iter) {
        const int64_t width = s.width;
        ds.rho = s.colptr<double>(Shard::C_RHO);
        ds.iters = s.colptr<int>(Shard::C_ITERS);
        ds.status = s.colptr<int>(Shard::C_STATUS);
        ds.fin = s.colptr<double>(Shard::C_FIN);
        double *z = s.colptr<double>(Shard::C_Z), *u = s.colptr<double>(Shard::C_U);
        double *xacc = s.colptr<double>(Shard::C_XACC);
        const double *par = s.par_batched ? s.colptr<double>(Shard::C_PAR) : s.par.p;
        // threads per problem of the prox kernel: enough CTAs x warps to fill the GPU at small widths
        int ch = width >= 32768 ? 4 : (width >= 8192 ? 8 : (width >= 2048 ? 16 : 32));
        int pw = ch == 16 ? 16 : (ch == 8 ? 24 : 32);
        if (const char *e = getenv("ADMMB_PROX_CH")) ch = atoi(e);
        if (const char *e = getenv("ADMMB_PROX_W")) pw = atoi(e);
        const unsigned gc = (unsigned)((width + 31) / 32);
        const int steps = std::min(chunk, op->max_iter - it);
        // every `refresh` iterations: x_R = exact FP64 x-update of the current (z, u), increment = 0 (drops the
        // rounding the tensor-core increments have accumulated in x_R)
        if (force_refresh || it >= next_refresh) {
            force_refresh = false;
            if (refresh > 0) next_refresh = (it - it0 < 512) ? it0 + 2 * std::max(it - it0, 32) : it + refresh;
            const unsigned gr = (unsigned)((width + 63) / 64);
            const double *s0w = s.colptr<double>(Shard::C_S0);
            if (s.has_c) k_tf32_final_x<true, 1><<<gr, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, zu_compact, nullptr, nullptr, s.ld_cur, nullptr, (int)width, nullptr, D.dscr.p, dscr_ld, xacc, s.ld_cur, ds.status);
            else k_tf32_final_x<false, 1><<<gr, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, zu_compact, nullptr, nullptr, s.ld_cur, nullptr, (int)width, nullptr, D.dscr.p, dscr_ld, xacc, s.ld_cur, ds.status);
            k_tf32_zero_running<<<dim3((unsigned)((width + 127) / 128), (unsigned)C.kpad), 128, 0, s.stream>>>(C.kpad, width, s.ld_cur, ds.status, C.bh, C.bl);
            CK(cudaGetLastError());
            s.launches += 2;
        }
        s.trace_active.push_back((int)width);
        auto launch_iter = [&](const DenseStep &d) {
            C.gemm(s.stream);
#define ADMMB_PROX_COND(CH, W)                                                                                           \
    k_prox_cond_tf32<CH, W><<<gc, dim3(32, CH), 0, s.stream>>>(s.nsplitblk, D.sblk.p, s.bdesc.p, width, s.ld_cur, par,  \
                                                               s.par_batched, op->alpha, C.Xr.p, xacc, z, u, zu_compact, C.bh, C.bl, d)
            if (ch == 4) { if (pw == 16) ADMMB_PROX_COND(4, 16); else if (pw == 24) ADMMB_PROX_COND(4, 24); else ADMMB_PROX_COND(4, 32); }
            else if (ch == 8) { if (pw == 16) ADMMB_PROX_COND(8, 16); else if (pw == 24) ADMMB_PROX_COND(8, 24); else ADMMB_PROX_COND(8, 32); }
            else if (ch == 16) { if (pw == 16) ADMMB_PROX_COND(16, 16); else ADMMB_PROX_COND(16, 32); }
            else ADMMB_PROX_COND(32, 32);
#undef ADMMB_PROX_COND
            s.launches += 2;
        };
        s.kernel_tic();
        auto tl0 = now();
        int k = 0;
        // long chunks replay a captured graph of GRAPH_ITERS iterations (2 launches each): the loop is launch-bound
        // at small widths; the iteration number then comes from a device counter.  The graph bakes in the pointers
        // of the current working set, so it is dropped at every repack.
        if (use_graphs && steps >= GRAPH_ITERS) {
            if (!gexec || gstale) {
                auto tg0 = now();
                cudaGraph_t g = nullptr;
                DenseStep dg = ds;
                dg.it_base = D.itbase.p;
                CK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
                for (int q = 1; q <= GRAPH_ITERS; ++q) { dg.it = q; launch_iter(dg); }
                k_add_int<<<1, 1, 0, s.stream>>>(D.itbase.p, GRAPH_ITERS);
                CK(cudaStreamEndCapture(s.stream, &g));
                s.launches -= 2 * GRAPH_ITERS;   // captured, not launched
                // same topology as before a repack, new pointers / grids: update the executable in place
                // (re-instantiating costs far more host time); fall back to a fresh instantiation if refused
                bool ok = false;
                if (gexec) {
                    cudaGraphExecUpdateResultInfo info;
                    ok = cudaGraphExecUpdate(gexec, g, &info) == cudaSuccess;
                    if (!ok) { (void)cudaGetLastError(); cudaGraphExecDestroy(gexec); gexec = nullptr; }
                }
                if (!ok) CK(cudaGraphInstantiate(&gexec, g, 0));
                cudaGraphDestroy(g);
                gstale = false;
                t_sec[1] += since(tg0);
            }
            k_set_int<<<1, 1, 0, s.stream>>>(D.itbase.p, it);
            for (; steps - k >= GRAPH_ITERS; k += GRAPH_ITERS, it += GRAPH_ITERS) {
                CK(cudaGraphLaunch(gexec, s.stream));
                s.launches += 2 * GRAPH_ITERS + 1;
            }
        }
        for (; k < steps; ++k) {
            ++it;
            ds.it = it;
            launch_iter(ds);
        }
        s.kernel_toc();
        t_sec[0] += since(tl0);
        auto ts0 = now();
        CK(cudaGetLastError());
        // who is still running?
        CK(cudaMemsetAsync(s.split_counts.p, 0, 2 * sizeof(int), s.stream));
        k_split<<<(unsigned)((width + 255) / 256), 256, 0, s.stream>>>(ds.status, (int)width, s.keep_list.p, s.fin_list.p,
                                                                      s.split_counts.p);
        ++s.launches;
        int cnt[2] = {0, 0};
        CK(cudaMemcpyAsync(cnt, s.split_counts.p, sizeof(cnt), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        t_sec[2] += since(ts0);
        if (op->chunk <= 0) {   // same launch-length policy as the Riccati path: <= ~2 % of the set finishing per check
            const int prev = chunk;
            if (cnt[1] == 0) chunk = std::min(chunk * 2, chunk_max);
            else chunk = (int)std::min<long long>(chunk_max, std::max<long long>(10, (long long)width * prev / (50LL * cnt[1])));
        }
        const bool last = cnt[0] == 0 || it >= op->max_iter;
        if (cnt[1] == 0) continue;
        // finished problems idle in the working set until at least 1/32 of it has finished (a repack costs a few
        // iterations' worth of traffic); a finished problem's columns are never written again, so retiring it later
        // is safe
        if (!last && (no_repack || (int64_t)cnt[1] * 32 < width)) continue;
        auto tf0 = now();
        {
            const unsigned gf = (unsigned)((cnt[1] + 63) / 64);
            const int *orig = s.cur_set < 0 ? nullptr : s.orig[s.cur_set].p;
            const double *s0w = s.colptr<double>(Shard::C_S0);
            if (tail) {   // leave d of the last right-hand side in the working set; k_output rebuilds x at download
                double *dw = s.colptr<double>(Shard::C_D);
                if (s.has_c) k_tf32_final_x<true, 2><<<gf, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, 1, C.bh, C.bl, s.ld_cur, s.fin_list.p, cnt[1], nullptr, dw, s.ld_cur, nullptr, 0, nullptr);
                else k_tf32_final_x<false, 2><<<gf, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, 1, C.bh, C.bl, s.ld_cur, s.fin_list.p, cnt[1], nullptr, dw, s.ld_cur, nullptr, 0, nullptr);
            } else {
                if (s.has_c) k_tf32_final_x<true, 0><<<gf, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, 0, C.bh, C.bl, s.ld_cur, s.fin_list.p, cnt[1], orig, D.dscr.p, dscr_ld, D.x.p, s.ld, nullptr);
                else k_tf32_final_x<false, 0><<<gf, 64, 0, s.stream>>>(N, s.fac.p, s.bdesc.p, s0w, z, u, 0, C.bh, C.bl, s.ld_cur, s.fin_list.p, cnt[1], orig, D.dscr.p, dscr_ld, D.x.p, s.ld, nullptr);
            }
            ++s.launches;
        }
        t_sec[3] += since(tf0);
        auto tr0 = now();
        s.repack(last ? 0 : cnt[0], cnt[1]);
        t_sec[4] += since(tr0);
        gstale = true;
        auto tb0 = now();
        if (!last) C.bind(s.colptr<float>(Shard::C_BH), s.colptr<float>(Shard::C_BL), s.ld_cur);
        t_sec[5] += since(tb0);
    }
    if (trace)
        fprintf(stderr, "[admmb trace] dense host ms: setup %.1f launch %.1f (graph build %.1f) split+sync %.1f final-x %.1f repack %.1f bind %.1f\n",
                t_setup, t_sec[0], t_sec[1], t_sec[2], t_sec[3], t_sec[4], t_sec[5]);
}

void dense_run(Shard &s, const admmb_opts *op)
{
    DenseState &D = s.dense;
    const int n = s.n, nb = s.nb;
    const unsigned gb = (unsigned)((s.batch + 127) / 128);
    // factor (shared model, one rho): Riccati factor is already in s.fac
    dense_build_factor(s, s.fac.p, s.has_c, D.M.p, D.S.p, D.mc.p);
    D.ready = true;
    const bool tf32 = op->precision == ADMMB_PREC_TF32;
    const int split = getenv("ADMMB_TF32_SINGLE") ? 1 : 3;
    CK(cudaMemsetAsync(D.x.p, 0, sizeof(double) * (size_t)n * s.ld, s.stream));
    // condensed form: only the split rows take part in the per-iteration GEMM (see dense_tf32.cuh)
    // (needs: no linear cost, states unsplit or not -- every CONTROL block split so the roll-out determines x)
    bool controls_split = true;
    for (int k = 0; k < s.N; ++k) controls_split = controls_split && (s.h_bdesc[3 * k + 2] & 0xff) != BLK_NONE;
    D.condensed = tf32 && !s.has_q && controls_split && getenv("ADMMB_NO_CONDENSED") == nullptr;
    if (D.condensed) {
        dense_run_condensed(s, op, split);
        return;
    }
    D.rt.alloc((size_t)s.n * s.ld);
    k_dense_rt_init<<<gb, 128, 0, s.stream>>>(nb, s.batch, s.ld, s.bdesc.p, s.z.p, s.u.p, s.rho.p,
                                              s.has_q ? s.q.p : nullptr, s.q_batched, D.rt.p);
    ++s.launches;
    if (tf32) {
        D.tf32.prepare(n, s.batch, s.ld, split, D.M.p, D.S.p, D.mc.p, s.s0.p, s.stream);
        dim3 g((unsigned)((s.ld + 127) / 128), (unsigned)n);
        k_tf32_split_rows<<<g, 128, 0, s.stream>>>(n, s.ld, D.rt.p, D.tf32.Bhi.p, split == 3 ? D.tf32.Blo.p : nullptr);
        s.launches += 3;
    }
    DenseStep ds{};
    ds.max_iter = op->max_iter;
    ds.reltol = op->reltol;
    ds.sqrtn_abs = sqrt((double)s.nsplit) * op->abstol;
    ds.rho = s.rho.p; ds.iters = s.iters.p; ds.status = s.status.p; ds.fin = s.fin.p; ds.running = D.running.p;
    const int chunk = op->chunk > 0 ? op->chunk : 25;
    dim3 gg((unsigned)((s.batch + DG_BN - 1) / DG_BN), (unsigned)((n + DG_BM - 1) / DG_BM));
    int running = 1;
    bool timing = false;
    for (int it = 1; it <= op->max_iter && running > 0; ++it) {
        if (!timing) { s.kernel_tic(); timing = true; }
        const bool check = (it % chunk) == 0 || it == op->max_iter;
        if (check) CK(cudaMemsetAsync(D.running.p, 0, sizeof(int), s.stream));
        ds.it = it;
        if (tf32) {
            dense_tf32_xupdate(s);
            k_prox_dual_residuals<true, true><<<gb, 128, 0, s.stream>>>(
                nb, s.batch, s.ld, s.bdesc.p, s.par.p, s.par_batched, nullptr, 0.0, op->alpha, D.tf32.X.p, s.z.p, s.u.p,
                nullptr, D.tf32.Bhi.p, D.tf32.split == 3 ? D.tf32.Blo.p : nullptr, D.x.p, s.has_q ? s.q.p : nullptr,
                s.q_batched, ds);
            s.launches += 2;
        } else {
            k_dense_xupdate_f64<<<gg, 256, 0, s.stream>>>(n, s.batch, s.ld, D.M.p, D.S.p, D.mc.p, s.s0.p, D.rt.p,
                                                          s.status.p, D.x.p);
            k_prox_dual_residuals<true, false><<<gb, 128, 0, s.stream>>>(
                nb, s.batch, s.ld, s.bdesc.p, s.par.p, s.par_batched, nullptr, 0.0, op->alpha, D.x.p, s.z.p, s.u.p,
                nullptr, D.rt.p, nullptr, nullptr, s.has_q ? s.q.p : nullptr, s.q_batched, ds);
            s.launches += 2;
        }
        CK(cudaGetLastError());
        if (check) {
            s.kernel_toc();
            timing = false;
            CK(cudaMemcpyAsync(&running, D.running.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
        }
    }
    if (timing) s.kernel_toc();
}

void dense_output(Shard &s, double *xo, double *zo, double *uo)
{
    DenseState &D = s.dense;
    const unsigned gb = (unsigned)((s.batch + 127) / 128);
    if (xo) CK(cudaMemcpyAsync(xo, D.x.p, sizeof(double) * (size_t)s.n * s.ld, cudaMemcpyDeviceToDevice, s.stream));
    if (zo || uo) {
        k_dense_output<<<gb, 128, 0, s.stream>>>(s.nb, s.batch, s.ld, s.bdesc.p, D.x.p, s.z.p, s.u.p, zo, uo);
        ++s.launches;
        CK(cudaGetLastError());
    }
}

}  // namespace
