// admm_b200.cu -- host side of libadmm_b200.so: the C ABI of include/admm_b200.h, device memory
// management, per-GPU shards and the launch loop of the persistent ADMM kernel.
//
// There is no reference implementation to mirror (/root/reference/README.md:1-2 is the whole of
// the reference); the behaviour implemented here is the one written down in oracle/admm_ocp.m
// (the MATLAB text BASELINE.json's north_star mandates) and SURVEY.md section 8(b).
// No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <new>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/admm_b200.h"
#include "host_util.cuh"
#include "kernels.cuh"
#include "iterate_launch_decl.cuh"
#include "dense.cuh"
#include "generators.cuh"
#include "scp.cuh"

using namespace admmb;

namespace {

thread_local std::string g_create_error;

struct Shard;
void dense_prepare(Shard &s, const admmb_opts *op);
void dense_run(Shard &s, const admmb_opts *op);
void dense_tail(Shard &s, const admmb_opts *op, int it0);
void dense_output(Shard &s, double *xo, double *zo, double *uo);
void dense_tf32_xupdate(Shard &s);
int dense_tf32_unit(Shard &s, int n, int64_t batch, size_t ld, const double *M, const double *S, const double *mc,
                    const double *s0, const double *rt, double *x, int split);

// ------------------------------------------------------------------------------------------------
// one GPU's share of a batch
// ------------------------------------------------------------------------------------------------
struct Shard {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = NUM_SMS_B200;
    int64_t launches = 0;
    std::vector<cudaEvent_t> kev;     // event pairs around the dominant kernel's launches
    size_t kev_used = 0;
    double kernel_ms = 0.0;
    int64_t kernel_launches = 0;
    void kernel_tic()
    {
        if (kev_used + 2 > kev.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); kev.push_back(e); }
        }
        CK(cudaEventRecord(kev[kev_used], stream));
    }
    void kernel_toc() { CK(cudaEventRecord(kev[kev_used + 1], stream)); kev_used += 2; }
    std::vector<int> trace_active;    // ADMMB_TRACE=1: active problems per launch
    void kernel_collect()   // after a stream sync
    {
        kernel_ms = 0.0;
        kernel_launches = (int64_t)(kev_used / 2);
        const bool trace = getenv("ADMMB_TRACE") != nullptr;
        for (size_t i = 0; i < kev_used; i += 2) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, kev[i], kev[i + 1]));
            kernel_ms += ms;
            if (trace && i / 2 < trace_active.size())
                fprintf(stderr, "[admmb trace] launch %zu: active %d, %.3f ms\n", i / 2, trace_active[i / 2], ms);
        }
        kev_used = 0;
        trace_active.clear();
    }

    // problem shape
    int N = 0, nb = 0, n = 0, nsplitblk = 0, rows_zu = 0, nsplit = 0;
    int64_t batch = 0, p_begin = 0;
    size_t ld = 0;
    bool dyn_batched = false, has_c = false, has_Q = false, has_R = false, has_q = false,
         q_batched = false, par_batched = false, has_z0 = false, has_u0 = false, has_rho0 = false;
    bool shared_factor = true, uploaded = false, ran = false;
    bool use_dense = false;
    bool tf32_tail = false;           // precision = tf32 with xupdate = auto: Riccati kernel while the working set is wide,
                                      // condensed incremental tensor-core pair once it is narrow (dense_tail)
    bool fast_pattern = false;   // stage states unsplit, every control split: prefetching kernel variant
    int kernel_variant = KV_AUTO;  // opts.kernel of the current run (ADMMB_KERNEL_*)
    int last_kernel = KV_THREAD;   // what the last launch ran
    bool decoupled = false;      // in-plane / cross-track structure proven on the factor: packed records
    bool time_invariant = false; // shared model with A_k, B_k bitwise equal for every stage (the warp-group kernel keeps them in registers)
    int max_iter_alloc = 0;
    bool hist_alloc = false;

    std::vector<int> h_bdesc, h_rowmap;
    DevBuf<double> rawA, rawB, rawc, rawQ, rawR, fac, fac_dec, s0, z, u, d, q, par, z0c, u0c, rho, rho0, usc, fin,
        hist, xo, zo, uo, stage;
    DevBuf<int> dec_flag, bdesc, rowmap, iters, status, fac_status, istage;
    // the working set: per-problem arrays that travel with the still-running problems (physical
    // compaction); set -1 = the home arrays themselves, 0/1 = dense ping-pong copies
    enum { C_Z, C_U, C_D, C_S0, C_RHO, C_USC, C_ITERS, C_STATUS, C_FIN, C_Q, C_PAR, C_FAC, C_FACDEC, C_RAWA, C_RAWB,
           C_RAWC, C_RAWQ, C_RAWR, C_XACC, C_BH, C_BL, C_COUNT };
    struct ColArray {
        void *home = nullptr;
        void *cur_override = nullptr;   // array created mid-run at the current pitch: stands in for work[cur_set] until the next repack
        size_t elem = 8;
        int rows = 0;          // 0: array absent / shared (does not travel)
        bool retire = false;   // written by the solver: finished problems copy it back to their home column
        DevBuf<unsigned char> work[2];
    };
    ColArray cols[C_COUNT];
    DevBuf<int> orig[2], keep_list, fin_list, split_counts, snap;
    int cur_set = -1;
    int64_t width = 0;
    int64_t n_real = 0;              // columns 0 .. n_real-1 of the working set are problems, the rest whole-warp padding (copies)
    size_t ld_cur = 0;
    template <typename T>
    T *colptr(int c) const
    {
        const ColArray &a = cols[c];
        if (a.cur_override) return (T *)a.cur_override;
        if (a.rows == 0 || cur_set < 0) return (T *)a.home;
        return (T *)a.work[cur_set].p;
    }
    void set_col(int c, void *home, size_t elem, int rows, bool retire)
    {
        cols[c].home = home; cols[c].elem = elem; cols[c].rows = home ? rows : 0; cols[c].retire = retire;
        cols[c].cur_override = nullptr;
    }
    void repack(int n_keep, int n_fin);
    bool pad_warps = getenv("ADMMB_NO_PAD") == nullptr;
    DevBuf<unsigned long long> counters;   // [0] refactor count, [1] converged, [2] sum iters, [3] max iters
    DenseState dense;

    void init(int dev)
    {
        device = dev;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        num_sms = prop.multiProcessorCount;
    }
    void destroy()
    {
        cudaSetDevice(device);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (auto e : kev) cudaEventDestroy(e);
        kev.clear();
        if (own_stream) cudaStreamDestroy(own_stream);
    }

    // host [cnt][R] -> device [R'][ld]
    void upload_rows(const double *host, int64_t cnt, int R, double *dst, const int *d_rowmap)
    {
        stage.alloc((size_t)cnt * R);
        CK(cudaMemcpyAsync(stage.p, host, sizeof(double) * (size_t)cnt * R, cudaMemcpyHostToDevice, stream));
        dim3 grid((unsigned)((cnt + 31) / 32), (unsigned)((R + 31) / 32)), block(32, 8);
        k_transpose_in<<<grid, block, 0, stream>>>(stage.p, cnt, R, dst, ld, d_rowmap);
        ++launches;
        CK(cudaGetLastError());
    }
    // device [R][ld] -> host [cnt][R]
    template <typename T>
    void download_rows(const T *src, int R, T *host, DevBuf<T> &stg)
    {
        stg.alloc((size_t)batch * R);
        dim3 grid((unsigned)((batch + 31) / 32), (unsigned)((R + 31) / 32)), block(32, 8);
        k_transpose_out<T><<<grid, block, 0, stream>>>(src, ld, R, batch, stg.p);
        ++launches;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host, stg.p, sizeof(T) * (size_t)batch * R, cudaMemcpyDeviceToHost, stream));
    }

    void upload(const admmb_problem *pb, const admmb_opts *op, int64_t begin, int64_t cnt, const admmb_generator *gen);
    void generate_model(const admmb_generator *gen, int64_t begin, int64_t cnt, int N_, size_t ld_, double *A, double *B,
                        DevBuf<double> &par_e, DevBuf<double> &par_th);
    DevBuf<double> gen_e, gen_th;
    DevBuf<double> pint_tab;          // parallel-in-time kernel: tables of the current shared factor (iterate_pint.cuh)
    DevBuf<double> wgpp_blk;          // streamed-record warp-group kernel: tile-blocked copy of the per-problem records (iterate_wg.cuh)
    bool pint_ready = false;
    void run(const admmb_opts *op, admmb_result *res);
    void download(admmb_result *res);
    void shift_warm_start(int k, const double *s0_new_host);
    // sequential convex programming on the resident batch (scp.cuh)
    bool scp_upload = false;          // set around upload(): per-problem model + affine term, produced on the device
    const int *run_active = nullptr;  // run(): problems with a zero entry keep the state of their last solve
    DevBuf<double> scp_xref, scp_step, scp_hist, scp_e, scp_th0, scp_theta;   // scp_theta [N][ld]: true anomaly at every stage start
    DevBuf<int> scp_active, scp_passes, scp_status, scp_count;
    DevBuf<long long> scp_iters;
    int scp_max_pass = 0;
    double scp_lin_ms = 0.0;
    void scp_linearise(const ScpConst &C, bool shoot);
    void scp_solve(const admmb_scp *sc, const admmb_opts *op, admmb_result *res);
    void scp_download(admmb_scp_result *out);
    template <bool FSH, bool FSMEM>
    void launch_iterate(const IterParams &P, bool adapt);
};

// the stage matrices computed on the device (csrc/generators.cuh) into the raw-model arrays
void Shard::generate_model(const admmb_generator *gen, int64_t begin, int64_t cnt, int N_, size_t ld_, double *A, double *B,
                           DevBuf<double> &par_e, DevBuf<double> &par_th)
{
    if (gen->kind == GEN_ELLIPTIC_ZOH) {
        par_e.alloc((size_t)cnt);
        par_th.alloc((size_t)cnt);
        CK(cudaMemcpyAsync(par_e.p, gen->e + begin, sizeof(double) * cnt, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(par_th.p, gen->theta0 + begin, sizeof(double) * cnt, cudaMemcpyHostToDevice, stream));
        dim3 grid((unsigned)((cnt + 127) / 128), 9);
        k_gen_elliptic<<<grid, 128, 0, stream>>>(par_e.p, par_th.p, cnt, N_, gen->T, gen->substeps > 0 ? gen->substeps : 8,
                                                 A, B, ld_);
    } else {
        k_gen_cw<<<1, 64, 0, stream>>>(gen->kind, N_, gen->T, gen->nmm != 0.0 ? gen->nmm : 1.0, A, B);
    }
    ++launches;
    CK(cudaGetLastError());
}

void Shard::upload(const admmb_problem *pb, const admmb_opts *op, int64_t begin, int64_t cnt, const admmb_generator *gen)
{
    CK(cudaSetDevice(device));
    uploaded = false;
    ran = false;
    N = pb->N;
    nb = 3 * N + 2;
    n = 9 * N + 6;
    batch = cnt;
    p_begin = begin;
    ld = round_up((size_t)cnt, 32);
    dyn_batched = scp_upload ? true : (gen ? gen->kind == GEN_ELLIPTIC_ZOH : pb->dyn_batched != 0);
    has_c = scp_upload || pb->c != nullptr;
    has_Q = pb->Q != nullptr;
    has_R = pb->R != nullptr;
    has_q = pb->q != nullptr;
    q_batched = has_q && pb->q_batched;
    par_batched = pb->par_batched != 0;
    has_z0 = pb->z0 != nullptr;
    has_u0 = pb->u0 != nullptr;
    has_rho0 = pb->rho0 != nullptr;

    // block descriptors: type | compact slot << 8
    h_bdesc.assign(nb, 0);
    h_rowmap.assign(n, -1);
    nsplitblk = 0;
    for (int b = 0; b < nb; ++b) {
        int t = pb->block_type[b];
        if (t != BLK_NONE) {
            for (int e = 0; e < 3; ++e) h_rowmap[3 * b + e] = 3 * nsplitblk + e;
            h_bdesc[b] = t | (nsplitblk << 8);
            ++nsplitblk;
        } else {
            h_bdesc[b] = t;
        }
    }
    rows_zu = 3 * nsplitblk;
    nsplit = rows_zu;
    use_dense = (op->xupdate == ADMMB_XUPDATE_DENSE);
    if (use_dense) {   // the dense path keeps full-width iterates (BLK_NONE rows stay zero)
        rows_zu = n;
        for (int i = 0; i < n; ++i) h_rowmap[i] = i;
    }
    fast_pattern = true;
    for (int k = 0; k < N; ++k)
        fast_pattern = fast_pattern && pb->block_type[3 * k] == BLK_NONE && pb->block_type[3 * k + 1] == BLK_NONE &&
                       pb->block_type[3 * k + 2] != BLK_NONE;
    bdesc.alloc(((size_t)nb + 3) / 4 * 4);   // padded: staged into smem with a 16-byte-granular bulk copy
    rowmap.alloc(n);
    CK(cudaMemsetAsync(bdesc.p, 0, sizeof(int) * (((size_t)nb + 3) / 4 * 4), stream));
    CK(cudaMemcpyAsync(bdesc.p, h_bdesc.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(rowmap.p, h_rowmap.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream));

    const bool has_P = has_Q || has_R;
    const bool per_rho = op->adapt_rho || has_rho0;
    shared_factor = !dyn_batched && (!has_P || !per_rho);
    {
        bool controls_split = true;
        for (int k = 0; k < N; ++k) controls_split = controls_split && pb->block_type[3 * k + 2] != BLK_NONE;
        tf32_tail = op->precision == ADMMB_PREC_TF32 && op->xupdate == ADMMB_XUPDATE_AUTO && shared_factor && !has_q &&
                    controls_split && !op->adapt_rho && !op->history && getenv("ADMMB_NO_TF32_TAIL") == nullptr;
    }

    time_invariant = !dyn_batched;
    for (int k = 1; k < N && time_invariant && !gen && !scp_upload; ++k)
        time_invariant = memcmp(pb->A, pb->A + (size_t)36 * k, 36 * sizeof(double)) == 0 &&
                         memcmp(pb->B, pb->B + (size_t)18 * k, 18 * sizeof(double)) == 0;

    // raw model
    const size_t md = dyn_batched ? ld : 1;
    auto up_model = [&](DevBuf<double> &buf, const double *host, int R) {
        if (!host) { buf.release(); return; }
        buf.alloc((size_t)R * md);
        if (dyn_batched) upload_rows(host + (size_t)begin * R, cnt, R, buf.p, nullptr);
        else CK(cudaMemcpyAsync(buf.p, host, sizeof(double) * R, cudaMemcpyHostToDevice, stream));
    };
    if (scp_upload) {   // the linearisation kernels fill these each pass
        rawA.alloc((size_t)36 * N * md);
        rawB.alloc((size_t)18 * N * md);
        rawc.alloc((size_t)6 * N * md);
        CK(cudaMemsetAsync(rawA.p, 0, sizeof(double) * 36 * N * md, stream));
        CK(cudaMemsetAsync(rawB.p, 0, sizeof(double) * 18 * N * md, stream));
        CK(cudaMemsetAsync(rawc.p, 0, sizeof(double) * 6 * N * md, stream));
    } else if (gen) {
        rawA.alloc((size_t)36 * N * md);
        rawB.alloc((size_t)18 * N * md);
        if (dyn_batched) {   // columns beyond the batch stay finite
            CK(cudaMemsetAsync(rawA.p, 0, sizeof(double) * 36 * N * md, stream));
            CK(cudaMemsetAsync(rawB.p, 0, sizeof(double) * 18 * N * md, stream));
        }
        generate_model(gen, begin, cnt, N, ld, rawA.p, rawB.p, gen_e, gen_th);
    } else {
        up_model(rawA, pb->A, 36 * N);
        up_model(rawB, pb->B, 18 * N);
    }
    if (!scp_upload) up_model(rawc, pb->c, 6 * N);
    up_model(rawQ, pb->Q, 36 * (N + 1));
    up_model(rawR, pb->R, 9 * N);

    s0.alloc(6 * ld);
    upload_rows(pb->s0 + (size_t)begin * 6, cnt, 6, s0.p, nullptr);
    if (has_q) {
        if (q_batched) { q.alloc((size_t)n * ld); upload_rows(pb->q + (size_t)begin * n, cnt, n, q.p, nullptr); }
        else { q.alloc(n); CK(cudaMemcpyAsync(q.p, pb->q, sizeof(double) * n, cudaMemcpyHostToDevice, stream)); }
    }
    if (par_batched) {
        par.alloc((size_t)8 * nb * ld);
        upload_rows(pb->block_par + (size_t)begin * 8 * nb, cnt, 8 * nb, par.p, nullptr);
    } else {
        par.alloc((size_t)8 * nb);
        CK(cudaMemcpyAsync(par.p, pb->block_par, sizeof(double) * 8 * nb, cudaMemcpyHostToDevice, stream));
    }
    if (has_z0) { z0c.alloc((size_t)rows_zu * ld); upload_rows(pb->z0 + (size_t)begin * n, cnt, n, z0c.p, rowmap.p); }
    if (has_u0) { u0c.alloc((size_t)rows_zu * ld); upload_rows(pb->u0 + (size_t)begin * n, cnt, n, u0c.p, rowmap.p); }
    if (has_rho0) {
        rho0.alloc(ld);
        CK(cudaMemcpyAsync(rho0.p, pb->rho0 + begin, sizeof(double) * cnt, cudaMemcpyHostToDevice, stream));
    }

    fac.alloc(shared_factor ? (size_t)FS * N : (size_t)FS * N * ld);
    z.alloc((size_t)rows_zu * ld);
    u.alloc((size_t)rows_zu * ld);
    d.alloc((size_t)3 * N * ld);
    rho.alloc(ld);
    usc.alloc(ld);
    fin.alloc(4 * ld);
    iters.alloc(ld);
    status.alloc(ld);
    fac_status.alloc(ld);
    keep_list.alloc(ld);
    fin_list.alloc(ld);
    snap.alloc(ld);
    split_counts.alloc(2);
    counters.alloc(4);
    CK(cudaMemsetAsync(fac_status.p, 0, sizeof(int) * ld, stream));
    CK(cudaMemsetAsync(d.p, 0, sizeof(double) * 3 * N * ld, stream));
    CK(cudaMemsetAsync(fin.p, 0, sizeof(double) * 4 * ld, stream));
    if (op->history) {
        const size_t need = (size_t)5 * op->max_iter * ld;
        if (need * sizeof(double) > ((size_t)64 << 30)) throw CudaFail{cudaErrorMemoryAllocation, "history too large"};
        hist.alloc(need);
    }
    max_iter_alloc = op->max_iter;
    hist_alloc = op->history != 0;
    CK(cudaStreamSynchronize(stream));
    uploaded = true;
}

// Widest working set the warp-group kernel takes when the choice is the library's, in PASSES (tiles per CTA).  Measured on
// B200 (scripts/variant_rates.py): a warp-group iteration costs ~0.2 us per stage and pass (9.9 us at N = 50), one problem
// per thread ~0.64 us per stage at any width up to ~16 k problems (32 us at N = 50), so the warp-group kernel wins up to
// three passes: 13,320 problems at N = 50 (30 problems per tile), 5,328 at N = 100 (12 per tile).  ADMMB_WG_WIDTH caps the
// width in problems instead (tuning knob; neither has any effect on results).
static bool wg_takes(int64_t n_active, int tile, int num_sms)
{
    static const int64_t w = getenv("ADMMB_WG_WIDTH") ? atoll(getenv("ADMMB_WG_WIDTH")) : -1;
    if (tile <= 0) return false;
    if (w >= 0) return n_active <= w;
    const int64_t tiles = (n_active + tile - 1) / tile;
    return (tiles + num_sms - 1) / num_sms <= 3;
}

// per-problem models: widest working set the streamed-record warp-group kernel takes when the choice is the library's, in
// passes (tiles per CTA), measured against k_admm_iterate_pptma (scripts/variant_rates.py cfg4); ADMMB_WGPP_PASSES overrides
static bool wgpp_takes(int64_t n_active, int tile, int num_sms)
{
    // measured (profiles/r2_cfg4_variant_rates.txt, N = 50, 24-problem tiles): 8,192 problems = 3 passes 73 vs 92 us per iteration,
    // 12,288 = 4 passes 98 vs 104, 16,384 = 5 passes 133 vs 118
    static const int64_t max_passes = getenv("ADMMB_WGPP_PASSES") ? atoll(getenv("ADMMB_WGPP_PASSES")) : 3;
    const int64_t tiles = (n_active + tile - 1) / tile;
    return (tiles + num_sms - 1) / num_sms <= max_passes;
}

template <bool FSH, bool FSMEM>
void Shard::launch_iterate(const IterParams &P, bool adapt)
{
    static const bool p2_default = getenv("ADMMB_P2") ? atoi(getenv("ADMMB_P2")) != 0 : true;
    IterLaunchCtx c{stream, num_sms, N, nb, par_batched, has_c, has_q, fast_pattern, decoupled, p2_default, kernel_variant,
                    rows_zu, device, time_invariant, pint_ready ? pint_tab.p : nullptr};
    last_kernel = KV_THREAD;
    if (FSH && FSMEM) {
        const bool wg = kernel_variant == KV_WG || (kernel_variant == KV_AUTO && wg_takes(P.n_active, iterate_wg_tile_width(c), num_sms));
        const bool tile = kernel_variant == KV_TILE;
        if (kernel_variant == KV_PINT && launch_iterate_pint(c, P, adapt)) last_kernel = KV_PINT;
        else if (wg && launch_iterate_wg(c, P, adapt)) last_kernel = KV_WG;
        else if (tile && launch_iterate_res(c, P, adapt)) last_kernel = KV_TILE;
        else launch_iterate_smem(c, P, adapt);
    }
    else if (FSH) launch_iterate_gshared(c, P, adapt);
    else {
        // per-problem models: the warp-group kernel with streamed records when pinned, or -- the library's choice -- on narrow
        // working sets, where one problem per thread is bound by a lone warp's instruction stream (52-70 us per iteration)
        const int tw = iterate_wgpp_tile_width(c, adapt && P.has_P);
        const bool wgpp = tw > 0 && (kernel_variant == KV_WG || (kernel_variant == KV_AUTO && wgpp_takes(P.n_active, tw, num_sms)));
        if (wgpp) {                         // room for the tile-blocked copy of the records (grows on first use only)
            wgpp_blk.alloc(wgpp_block_doubles(c, P.n_active, tw));
            c.wgpp_blk = wgpp_blk.p;
        }
        if (wgpp && launch_iterate_wgpp(c, P, adapt)) last_kernel = KV_WG;
        else if (!launch_iterate_pptma(c, P, adapt)) launch_iterate_pp(c, P, adapt);
    }
    ++launches;
}

// retire the finished columns of the working set to their home columns and move the still-running
// ones into a dense, 32-aligned prefix of the other ping-pong buffer
void Shard::repack(int n_keep, int n_fin)
{
    const unsigned gf = (unsigned)((n_fin + 127) / 128);
    auto rows_grid = [](int rows) { return (unsigned)std::min(rows, 64); };
    if (cur_set >= 0 && n_fin > 0) {
        for (int c = 0; c < C_COUNT; ++c) {
            ColArray &a = cols[c];
            if (a.rows == 0 || !a.retire) continue;
            dim3 grid(gf, rows_grid(a.rows));
            const int *skip = (c == C_Z || c == C_U || c == C_D) ? snap.p : nullptr;
            if (a.elem == 8)
                k_scatter_cols<double><<<grid, 128, 0, stream>>>((const double *)a.work[cur_set].p, ld_cur, a.rows,
                                                                 fin_list.p, n_fin, orig[cur_set].p, (double *)a.home, ld, skip);
            else
                k_scatter_cols<int><<<grid, 128, 0, stream>>>((const int *)a.work[cur_set].p, ld_cur, a.rows, fin_list.p,
                                                              n_fin, orig[cur_set].p, (int *)a.home, ld);
            ++launches;
        }
        CK(cudaGetLastError());
    }
    CK(cudaMemsetAsync(snap.p, 0, sizeof(int) * ld, stream));   // the next working set starts without snapshots
    if (n_keep == 0) { width = 0; n_real = 0; return; }
    n_real = n_keep;
    // Pad the working set to whole warps with COPIES of its last running problem (own columns, same home column):
    // a launch in which one warp has idle lanes runs ~25 % slower on a narrow working set (measured: 8,160 problems
    // 1.62 ms per 50 iterations, 8,161 or 8,191 problems 2.02-2.04 ms).  The copies evolve identically to their
    // source and retire to the same home column with the same values.
    if (pad_warps && (n_keep & 31)) {
        const int n_pad = (int)round_up((size_t)n_keep, 32);
        k_pad_list<<<1, 32, 0, stream>>>(keep_list.p, n_keep, n_pad);
        ++launches;
        n_keep = n_pad;
    }
    const unsigned gk2 = (unsigned)((n_keep + 127) / 128);
    const int nxt = cur_set == 0 ? 1 : 0;
    const size_t ld_new = round_up((size_t)n_keep, 32);
    orig[nxt].alloc(ld_new);
    k_compose_orig<<<gk2, 128, 0, stream>>>(cur_set < 0 ? nullptr : orig[cur_set].p, keep_list.p, n_keep, orig[nxt].p);
    ++launches;
    for (int c = 0; c < C_COUNT; ++c) {
        ColArray &a = cols[c];
        if (a.rows == 0) continue;
        a.work[nxt].alloc((size_t)a.rows * ld_new * a.elem);
        const void *src = a.cur_override ? a.cur_override : (cur_set < 0 ? a.home : (const void *)a.work[cur_set].p);
        a.cur_override = nullptr;
        dim3 grid(gk2, rows_grid(a.rows));
        if (a.elem == 8)
            k_gather_cols<double><<<grid, 128, 0, stream>>>((const double *)src, ld_cur, a.rows, keep_list.p, n_keep,
                                                            (double *)a.work[nxt].p, ld_new);
        else
            k_gather_cols<int><<<grid, 128, 0, stream>>>((const int *)src, ld_cur, a.rows, keep_list.p, n_keep,
                                                         (int *)a.work[nxt].p, ld_new);
        ++launches;
    }
    CK(cudaGetLastError());
    cur_set = nxt;
    width = n_keep;
    ld_cur = ld_new;
}

// Receding-horizon warm start: split block `slot` of the next solve starts from the block `src[slot]`
// of the solution just computed (the same block k stages later; -1: beyond the horizon, start from zero).  u is stored
// with a pending scale after a rho change, so the scale is applied here.
__global__ void k_shift_warm(int64_t batch, size_t ld, int nslots, const int *src, const double *z, const double *u,
                             const double *usc, double *z0, double *u0)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const double sc = usc[p];
    for (int slot = blockIdx.y; slot < nslots; slot += gridDim.y) {
        const int from = src[slot];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const size_t r = (size_t)(3 * slot + e) * ld + p;
            z0[r] = from >= 0 ? z[(size_t)(3 * from + e) * ld + p] : 0.0;
            u0[r] = from >= 0 ? u[(size_t)(3 * from + e) * ld + p] * sc : 0.0;
        }
    }
}

// The solved batch stays on the device; its (z, u), shifted by k stages, become the warm start of the next solve and the
// initial states are replaced by s0_new (host, [6 x batch]) or, when that is null, by the solution's own state at stage k.
void Shard::shift_warm_start(int k, const double *s0_new_host)
{
    CK(cudaSetDevice(device));
    std::vector<int> src(nsplitblk, -1);
    for (int b = 0; b < nb; ++b) {
        if ((h_bdesc[b] & 0xff) == BLK_NONE) continue;
        const int slot = h_bdesc[b] >> 8;
        if (b >= 3 * N) { src[slot] = slot; continue; }          // terminal blocks keep their own iterate
        const int from = b + 3 * k;
        if (from < 3 * N && (h_bdesc[from] & 0xff) != BLK_NONE) src[slot] = h_bdesc[from] >> 8;
    }
    istage.alloc(nsplitblk);
    CK(cudaMemcpyAsync(istage.p, src.data(), sizeof(int) * nsplitblk, cudaMemcpyHostToDevice, stream));
    z0c.alloc((size_t)rows_zu * ld);
    u0c.alloc((size_t)rows_zu * ld);
    if (!s0_new_host) {
        // the solution's state at stage k: rows 9k .. 9k+5 of x (rebuilt from d exactly as for download)
        xo.alloc((size_t)n * ld);
        const unsigned gb = (unsigned)((batch + 127) / 128);
        if (shared_factor) {
            if (has_c) k_output<true, true><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, xo.p, nullptr, nullptr);
            else k_output<true, false><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, xo.p, nullptr, nullptr);
        } else {
            if (has_c) k_output<false, true><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, xo.p, nullptr, nullptr);
            else k_output<false, false><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, xo.p, nullptr, nullptr);
        }
        ++launches;
        CK(cudaGetLastError());
    }
    dim3 grid((unsigned)((batch + 127) / 128), (unsigned)std::min(nsplitblk, 64));
    k_shift_warm<<<grid, 128, 0, stream>>>(batch, ld, nsplitblk, istage.p, z.p, u.p, usc.p, z0c.p, u0c.p);
    ++launches;
    CK(cudaGetLastError());
    if (s0_new_host) upload_rows(s0_new_host + (size_t)p_begin * 6, batch, 6, s0.p, nullptr);
    else
        for (int i = 0; i < 6; ++i)
            CK(cudaMemcpyAsync(s0.p + (size_t)i * ld, xo.p + (size_t)(9 * k + i) * ld, sizeof(double) * ld,
                               cudaMemcpyDeviceToDevice, stream));
    has_z0 = has_u0 = true;
    CK(cudaStreamSynchronize(stream));       // `src` and the staging buffer are reused by the caller
}

// ------------------------------------------------------------------------------------------------
// sequential convex programming on the resident batch (kernels: scp.cuh, oracle: oracle/scp_ocp.py)
// ------------------------------------------------------------------------------------------------
void Shard::scp_linearise(const ScpConst &C, bool shoot)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, stream));
    const unsigned gb = (unsigned)((batch + 127) / 128);
    if (shoot) {
        k_scp_shoot<<<gb, 128, 0, stream>>>(C, batch, N, ld, s0.p, scp_xref.p, rawA.p, rawB.p, rawc.p, scp_e.p, scp_theta.p);
    } else {
        dim3 grid(gb, (unsigned)N);
        k_scp_linearise<<<grid, 128, 0, stream>>>(C, batch, N, ld, scp_active.p, scp_xref.p, rawA.p, rawB.p, rawc.p, scp_e.p,
                                                  scp_theta.p);
    }
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(b, stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (e != cudaSuccess) throw CudaFail{e, "scp_linearise"};
    scp_lin_ms += ms;
}

void Shard::scp_solve(const admmb_scp *sc, const admmb_opts *op, admmb_result *res)
{
    CK(cudaSetDevice(device));
    ScpConst C;
    const double nmm = sc->nmm != 0.0 ? sc->nmm : 1.0;
    C.substeps = sc->substeps > 0 ? sc->substeps : 8;
    C.R0 = sc->R0;
    C.twoR0 = 2.0 * sc->R0;
    C.R0sq = sc->R0 * sc->R0;
    C.n2 = nmm * nmm;
    C.tn = 2.0 * nmm;
    C.dt = sc->T / (double)C.substeps;
    C.hdt = 0.5 * C.dt;
    C.dt6 = C.dt / 6.0;
    C.impulsive = sc->control == ADMMB_SCP_CTRL_IMPULSIVE;
    C.elliptic = sc->model == ADMMB_SCP_NL_ELLIPTIC;
    if (C.elliptic) {   // the orbit's own clock: true anomaly at the start of every stage, once per solve
        scp_e.alloc(ld);
        scp_th0.alloc(ld);
        scp_theta.alloc((size_t)N * ld);
        CK(cudaMemcpyAsync(scp_e.p, sc->e + p_begin, sizeof(double) * batch, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(scp_th0.p, sc->theta0 + p_begin, sizeof(double) * batch, cudaMemcpyHostToDevice, stream));
        k_scp_theta<<<(unsigned)((batch + 127) / 128), 128, 0, stream>>>(batch, N, ld, C.dt, C.hdt, C.dt6, C.substeps, scp_e.p,
                                                                         scp_th0.p, scp_theta.p);
        ++launches;
        CK(cudaGetLastError());
    }
    scp_max_pass = sc->max_pass;
    scp_lin_ms = 0.0;
    scp_xref.alloc((size_t)n * ld);
    scp_active.alloc(ld);
    scp_passes.alloc(ld);
    scp_status.alloc(ld);
    scp_step.alloc(ld);
    scp_iters.alloc(ld);
    scp_hist.alloc((size_t)scp_max_pass * ld);
    scp_count.alloc(1);
    xo.alloc((size_t)n * ld);
    CK(cudaMemsetAsync(scp_xref.p, 0, sizeof(double) * n * ld, stream));           // first reference: free drift (a = 0)
    CK(cudaMemsetAsync(scp_hist.p, 0xFF, sizeof(double) * (size_t)scp_max_pass * ld, stream));   // NaN after a problem's exit
    const unsigned gb = (unsigned)((batch + 127) / 128);
    k_scp_init<<<(unsigned)((ld + 127) / 128), 128, 0, stream>>>(batch, ld, scp_active.p, scp_passes.p, scp_status.p,
                                                                 scp_step.p, scp_iters.p);
    ++launches;
    CK(cudaGetLastError());
    struct ActiveGuard { Shard &s; ~ActiveGuard() { s.run_active = nullptr; } } guard{*this};
    auto t0 = std::chrono::steady_clock::now();
    double kms = 0.0;
    int64_t kl = 0;
    for (int pass = 1; pass <= scp_max_pass; ++pass) {
        scp_linearise(C, pass == 1);
        if (pass > 1) {   // warm start: the previous pass's (z, u); converged problems are left alone
            z0c.alloc((size_t)rows_zu * ld);
            u0c.alloc((size_t)rows_zu * ld);
            dim3 grid(gb, (unsigned)std::min(rows_zu, 64));
            k_scp_warm<<<grid, 128, 0, stream>>>(batch, ld, rows_zu, z.p, u.p, usc.p, z0c.p, u0c.p);
            ++launches;
            CK(cudaGetLastError());
            has_z0 = has_u0 = true;
            run_active = scp_active.p;
        }
        admmb_result r;
        memset(&r, 0, sizeof(r));
        run(op, &r);
        kms += r.kernel_ms;
        kl += r.kernel_launches;
        k_output<false, true><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, xo.p,
                                                      nullptr, nullptr);
        CK(cudaMemsetAsync(scp_count.p, 0, sizeof(int), stream));
        k_scp_step<<<gb, 128, 0, stream>>>(batch, n, ld, pass, sc->tol_abs, sc->tol_rel, xo.p, scp_xref.p, iters.p, status.p,
                                           scp_active.p, scp_passes.p, scp_status.p, scp_step.p, scp_iters.p, scp_hist.p,
                                           scp_count.p);
        launches += 2;
        CK(cudaGetLastError());
        int still = 0;
        CK(cudaMemcpyAsync(&still, scp_count.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (still == 0) break;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (res) {
        res->kernel_ms = kms;
        res->kernel_launches = kl;
        res->device_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();   // every pass ends with a stream sync
    }
    ran = true;
}

void Shard::scp_download(admmb_scp_result *out)
{
    CK(cudaSetDevice(device));
    const size_t pb = (size_t)p_begin;
    std::vector<int> hp(batch), hs(batch);
    std::vector<long long> hi(batch);
    CK(cudaMemcpyAsync(hp.data(), scp_passes.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(hs.data(), scp_status.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(hi.data(), scp_iters.p, sizeof(long long) * batch, cudaMemcpyDeviceToHost, stream));
    if (out->step) CK(cudaMemcpyAsync(out->step + pb, scp_step.p, sizeof(double) * batch, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int64_t i = 0; i < batch; ++i) {
        out->stats[0] += hs[i] == 0;
        out->stats[1] += hi[i];
        out->stats[2] = std::max<int64_t>(out->stats[2], hp[i]);
        out->stats[3] += hp[i];
        if (out->passes) out->passes[pb + i] = hp[i];
        if (out->scp_status) out->scp_status[pb + i] = hs[i];
        if (out->iters_total) out->iters_total[pb + i] = hi[i];
    }
    if (out->hist_step) {
        download_rows<double>(scp_hist.p, scp_max_pass, out->hist_step + pb * scp_max_pass, stage);
        CK(cudaStreamSynchronize(stream));
    }
}

__global__ void k_stats(int64_t batch, const int *iters, const int *status, unsigned long long *counters)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long conv = 0, it = 0;
    if (p < batch) { conv = status[p] == ST_CONVERGED; it = (unsigned long long)iters[p]; }
    unsigned long long mx = it;
    for (int o = 16; o > 0; o >>= 1) {
        conv += __shfl_xor_sync(0xffffffffu, conv, o);
        it += __shfl_xor_sync(0xffffffffu, it, o);
        unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counters + 1, conv);
        atomicAdd(counters + 2, it);
        atomicMax(counters + 3, mx);
    }
}

void Shard::run(const admmb_opts *op, admmb_result *res)
{
    CK(cudaSetDevice(device));
    const bool has_P = has_Q || has_R;
    const bool adapt = op->adapt_rho != 0;
    kernel_variant = op->kernel;
    const unsigned gb = (unsigned)((batch + 127) / 128);
    CK(cudaEventRecord(ev0, stream));
    CK(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * 4, stream));
    CK(cudaMemsetAsync(fac_status.p, 0, sizeof(int) * ld, stream));
    // history entries of iterations a problem never runs are NaN (all-ones bit pattern), as in the oracle: callers of
    // the C ABI and the MEX gateway get them as they are
    if (op->history && hist_alloc) CK(cudaMemsetAsync(hist.p, 0xFF, sizeof(double) * 5 * (size_t)max_iter_alloc * ld, stream));

    // ---- factorisation (row a1), once per rho
    if (shared_factor) {
        k_riccati_factor<<<1, 32, 0, stream>>>(N, 1, 0, 0, rawA.p, rawB.p, rawc.p, rawQ.p, rawR.p, ld, nullptr,
                                                op->rho, bdesc.p, fac.p, nullptr);
    } else {
        k_riccati_factor<<<gb, 128, 0, stream>>>(N, batch, dyn_batched ? 1 : 0, 1, rawA.p, rawB.p, rawc.p, rawQ.p,
                                                 rawR.p, ld, has_rho0 ? rho0.p : nullptr, op->rho, bdesc.p, fac.p,
                                                 fac_status.p);
    }
    ++launches;
    CK(cudaGetLastError());
    decoupled = false;
    pint_ready = false;
    if (fast_pattern && !use_dense && getenv("ADMMB_NO_DECOUPLED") == nullptr) {
        const int64_t nf = shared_factor ? 1 : batch;
        dec_flag.alloc(1);
        CK(cudaMemsetAsync(dec_flag.p, 0, sizeof(int), stream));
        k_check_decoupled<<<(unsigned)((nf + 127) / 128), 128, 0, stream>>>(N, nf, shared_factor ? 0 : 1, fac.p, ld, dec_flag.p);
        ++launches;
        int flag = 1;
        CK(cudaMemcpyAsync(&flag, dec_flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (flag == 0) {
            decoupled = true;
            fac_dec.alloc(shared_factor ? (size_t)FD * N : (size_t)FD * N * ld);
            k_pack_decoupled<<<(unsigned)((nf + 127) / 128), 128, 0, stream>>>(N, nf, shared_factor ? 0 : 1, fac.p, ld, fac_dec.p);
            ++launches;
            CK(cudaGetLastError());
            if (shared_factor && op->kernel == KV_PINT) {   // opt-in parallel-in-time kernel: tables of this factor
                pint_tab.alloc(pint_table_size(N));
                launch_pint_pack(stream, N, fac_dec.p, pint_tab.p);
                ++launches;
                pint_ready = true;
            }
        }
    }
    k_reset<<<gb, 128, 0, stream>>>(batch, ld, rows_zu, z.p, u.p, has_z0 ? z0c.p : nullptr,
                                    has_u0 ? u0c.p : nullptr, rho.p, has_rho0 ? rho0.p : nullptr, op->rho, usc.p,
                                    iters.p, status.p, fac_status.p, run_active);
    ++launches;
    CK(cudaGetLastError());

    if (use_dense) {
        dense_run(*this, op);
    } else {
        IterParams P;
        memset(&P, 0, sizeof(P));
        P.N = N; P.nb = nb; P.n = n;
        P.raw_batched = dyn_batched;
        P.q_batched = q_batched;
        P.bdesc = bdesc.p; P.par_batched = par_batched;
        P.hist = op->history ? hist.p : nullptr;
        P.hist_stride = (size_t)op->max_iter * ld;
        P.hist_ld = ld;
        P.refac_count = counters.p;
        P.alpha = op->alpha; P.oma = 1.0 - op->alpha; P.reltol = op->reltol;
        P.sqrtn_abs = sqrt((double)nsplit) * op->abstol;
        P.mu = op->adapt_mu; P.tau = op->adapt_tau; P.inv_tau = 1.0 / op->adapt_tau;
        P.adapt = adapt; P.every = op->adapt_every > 0 ? op->adapt_every : 1; P.until = op->adapt_until;
        P.max_iter = op->max_iter; P.has_P = has_P;
        int chunk = op->chunk > 0 ? op->chunk : 50;
        if (adapt && P.every < chunk && P.every > 0) chunk = (chunk / P.every) * P.every;   // keep launches aligned
        if (chunk < 1) chunk = 1;
        P.chunk = chunk;
        const bool fsmem = shared_factor &&
                           sizeof(double) * (decoupled ? FD : FS) * N + sizeof(double) * 8 * nb + 4 * nb + 64 <= 200 * 1024;

        // which per-problem arrays travel with the working set
        const bool refac = !shared_factor && has_P && adapt;
        for (int c = 0; c < C_COUNT; ++c) set_col(c, nullptr, 8, 0, false);   // incl. what a previous run's tail registered
        set_col(C_Z, z.p, 8, rows_zu, true);
        set_col(C_U, u.p, 8, rows_zu, true);
        set_col(C_D, d.p, 8, 3 * N, true);
        set_col(C_S0, s0.p, 8, 6, false);
        set_col(C_RHO, rho.p, 8, 1, true);
        set_col(C_USC, usc.p, 8, 1, true);
        set_col(C_ITERS, iters.p, 4, 1, true);
        set_col(C_STATUS, status.p, 4, 1, true);
        set_col(C_FIN, fin.p, 8, 4, true);
        set_col(C_Q, (has_q && q_batched) ? q.p : nullptr, 8, n, false);
        set_col(C_PAR, par_batched ? par.p : nullptr, 8, 8 * nb, false);
        set_col(C_FAC, shared_factor ? nullptr : fac.p, 8, FS * N, refac);
        set_col(C_FACDEC, (!shared_factor && decoupled) ? fac_dec.p : nullptr, 8, FD * N, false);
        set_col(C_RAWA, (refac && dyn_batched) ? rawA.p : nullptr, 8, 36 * N, false);
        set_col(C_RAWB, (refac && dyn_batched) ? rawB.p : nullptr, 8, 18 * N, false);
        set_col(C_RAWC, (refac && dyn_batched && has_c) ? rawc.p : nullptr, 8, 6 * N, false);
        set_col(C_RAWQ, (refac && dyn_batched && has_Q) ? rawQ.p : nullptr, 8, 36 * (N + 1), false);
        set_col(C_RAWR, (refac && dyn_batched && has_R) ? rawR.p : nullptr, 8, 9 * N, false);
        cur_set = -1;
        width = batch;
        n_real = batch;
        ld_cur = ld;
        const bool no_repack = getenv("ADMMB_NO_REPACK") != nullptr;
        const bool zombies_enabled = getenv("ADMMB_NO_ZOMBIES") == nullptr;
        CK(cudaMemsetAsync(snap.p, 0, sizeof(int) * ld, stream));

        // launches get longer once few problems finish per launch: the host round trip (split, count
        // read-back, repack) then costs relatively less and nothing is lost in early-exit granularity
        const int chunk_max = op->chunk > 0 ? op->chunk : 400;
        // below this many running problems one warp's serial sweep bounds the Riccati kernel (~33 us per iteration)
        // and the GEMM + prox pair is faster (32 us at 8,192, 21 us at 1,024: DESIGN 6.3)
        const int64_t tail_width = op->tf32_switch > 0 ? op->tf32_switch : (op->tf32_switch < 0 ? 0 : 8192);   // opts.tf32_switch
        int wg_tile = 0;
        {
            IterLaunchCtx c{stream, num_sms, N, nb, par_batched, has_c, has_q, fast_pattern, decoupled, false, kernel_variant,
                            rows_zu, device, time_invariant, nullptr};
            if (shared_factor && fsmem) wg_tile = kernel_variant == KV_PINT && pint_ready ? iterate_pint_tile_width(c) : iterate_wg_tile_width(c);
            else if (!shared_factor) wg_tile = iterate_wgpp_tile_width(c, adapt && has_P);
        }
        int done_iters = 0;
        while (width > 0 && done_iters < op->max_iter && !(tf32_tail && width <= tail_width)) {
            P.chunk = chunk;
            P.ld = ld_cur;
            P.n_active = (int)width;
            P.n_real = (int)n_real;
            P.orig = cur_set < 0 ? nullptr : orig[cur_set].p;
            const bool zomb = cur_set >= 0 && zombies_enabled;
            P.z_home = zomb ? z.p : nullptr; P.u_home = zomb ? u.p : nullptr; P.d_home = zomb ? d.p : nullptr;
            P.home_ld = ld; P.rows_zu = rows_zu; P.snap = snap.p;
            P.fac = shared_factor ? fac.p : colptr<double>(C_FAC);
            P.fac_rw = colptr<double>(C_FAC);
            P.fac_dec = (shared_factor || !decoupled) ? fac_dec.p : colptr<double>(C_FACDEC);
            P.fac_dec_rw = colptr<double>(C_FACDEC);
            // shared models keep their single raw copy; per-problem raw models travel only when refactoring
            P.rawA = cols[C_RAWA].rows ? colptr<double>(C_RAWA) : rawA.p;
            P.rawB = cols[C_RAWB].rows ? colptr<double>(C_RAWB) : rawB.p;
            P.rawc = cols[C_RAWC].rows ? colptr<double>(C_RAWC) : rawc.p;
            P.rawQ = cols[C_RAWQ].rows ? colptr<double>(C_RAWQ) : rawQ.p;
            P.rawR = cols[C_RAWR].rows ? colptr<double>(C_RAWR) : rawR.p;
            P.s0 = colptr<double>(C_S0); P.z = colptr<double>(C_Z); P.u = colptr<double>(C_U); P.d = colptr<double>(C_D);
            P.q = has_q ? (q_batched ? colptr<double>(C_Q) : q.p) : nullptr;
            P.par = par_batched ? colptr<double>(C_PAR) : par.p;
            P.rho = colptr<double>(C_RHO); P.usc = colptr<double>(C_USC);
            P.iters = colptr<int>(C_ITERS); P.status = colptr<int>(C_STATUS); P.fin = colptr<double>(C_FIN);
            trace_active.push_back((int)width);
            kernel_tic();
            if (shared_factor) { if (fsmem) launch_iterate<true, true>(P, adapt); else launch_iterate<true, false>(P, adapt); }
            else launch_iterate<false, false>(P, adapt);
            kernel_toc();
            done_iters += chunk;
            // who is still running?
            CK(cudaMemsetAsync(split_counts.p, 0, 2 * sizeof(int), stream));
            // only the real columns are listed: the padding copies of the last problem (repack) end with this working set --
            // their source column carries the same state -- so they never pile up and never count in the statistics
            k_split<<<(unsigned)((n_real + 255) / 256), 256, 0, stream>>>(P.status, (int)n_real, keep_list.p, fin_list.p,
                                                                        split_counts.p);
            ++launches;
            CK(cudaGetLastError());
            int cnt[2] = {0, 0};
            CK(cudaMemcpyAsync(cnt, split_counts.p, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            // The warp-group kernel holds a whole tile of problems per CTA for the launch: finished problems idle in their
            // lanes at no cost to the others, and a launch's time only depends on the number of PASSES (tiles per CTA).  So
            // the working set is repacked only when that drops a pass, and a single-pass working set runs in long launches
            // (a CTA leaves as soon as all its problems are done).
            bool wg_skip = false;
            if ((last_kernel == KV_WG || last_kernel == KV_PINT) && wg_tile > 0 && cnt[0] > 0 &&
                (kernel_variant == KV_AUTO || kernel_variant == KV_PINT)) {
                auto passes = [&](int64_t w) { const int64_t tiles = (w + wg_tile - 1) / wg_tile; return (tiles + num_sms - 1) / num_sms; };
                const int64_t now = passes(width), then = passes((int64_t)round_up((size_t)cnt[0], 32));
                wg_skip = then >= now;
                if (op->chunk <= 0) chunk = now <= 1 ? 1000 : 200;
            } else if (op->chunk <= 0) {
                // adapt the launch length: finished problems idle until the launch ends, so aim at <= 1 % of the
                // working set finishing per launch, estimated from the finish rate of the launch just done
                const int prev = chunk;
                if (cnt[1] == 0) chunk = std::min(chunk * 2, chunk_max);
                else chunk = (int)std::min<long long>(chunk_max, std::max<long long>(50, (long long)width * prev / (100LL * cnt[1])));
            }
            if (adapt && P.every > 0 && chunk > P.every) chunk = (chunk / P.every) * P.every;
            if (cnt[0] == (int)n_real) continue;                // nobody finished in this launch
            if (wg_skip) continue;                              // warp-group kernel: no pass to gain from a repack
            if (no_repack && cnt[0] > 0 && cur_set < 0) continue;   // debug: finished lanes just idle
            repack(cnt[0], cnt[1]);
            if (tf32_tail && width > 0 && width <= tail_width) break;
        }
        // precision = tf32, xupdate = auto: the narrow remainder continues on the tensor-core pair
        if (tf32_tail && width > 0 && width <= tail_width && done_iters < op->max_iter) dense_tail(*this, op, done_iters);
        if (width > 0 && cur_set >= 0) {   // max_iter reached between checks: everything left is final
            CK(cudaMemsetAsync(split_counts.p, 0, 2 * sizeof(int), stream));
            k_iota<<<(unsigned)((n_real + 127) / 128), 128, 0, stream>>>(fin_list.p, (int)n_real);
            ++launches;
            repack(0, (int)n_real);
        }
    }
    k_stats<<<gb, 128, 0, stream>>>(batch, iters.p, status.p, counters.p);
    ++launches;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev1, stream));
    unsigned long long hc[4];
    CK(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    kernel_collect();
    if (res) {
        res->kernel_ms = kernel_ms;
        res->kernel_launches = kernel_launches;
        res->stats[0] = (int64_t)hc[1];
        res->stats[1] = (int64_t)hc[2];
        res->stats[2] = (int64_t)hc[3];
        res->stats[3] = (int64_t)hc[0];
        res->device_ms = ms;
    }
    ran = true;
}

void Shard::download(admmb_result *res)
{
    CK(cudaSetDevice(device));
    const size_t pb = (size_t)p_begin;
    if (res->x || res->z || res->u) {
        if (res->x) xo.alloc((size_t)n * ld);
        if (res->z) zo.alloc((size_t)n * ld);
        if (res->u) uo.alloc((size_t)n * ld);
        const unsigned gb = (unsigned)((batch + 127) / 128);
        if (use_dense) {
            dense_output(*this, res->x ? xo.p : nullptr, res->z ? zo.p : nullptr, res->u ? uo.p : nullptr);
        } else if (shared_factor) {
            if (has_c) k_output<true, true><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, res->x ? xo.p : nullptr, res->z ? zo.p : nullptr, res->u ? uo.p : nullptr);
            else k_output<true, false><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, res->x ? xo.p : nullptr, res->z ? zo.p : nullptr, res->u ? uo.p : nullptr);
        } else {
            if (has_c) k_output<false, true><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, res->x ? xo.p : nullptr, res->z ? zo.p : nullptr, res->u ? uo.p : nullptr);
            else k_output<false, false><<<gb, 128, 0, stream>>>(N, batch, ld, fac.p, s0.p, d.p, z.p, u.p, usc.p, bdesc.p, iters.p, res->x ? xo.p : nullptr, res->z ? zo.p : nullptr, res->u ? uo.p : nullptr);
        }
        ++launches;
        CK(cudaGetLastError());
        // one staging buffer per output so the three D2H copies can be queued back to back
        static_assert(sizeof(double) == 8, "");
        if (res->x) { download_rows<double>(xo.p, n, res->x + pb * n, stage); CK(cudaStreamSynchronize(stream)); }
        if (res->z) { download_rows<double>(zo.p, n, res->z + pb * n, stage); CK(cudaStreamSynchronize(stream)); }
        if (res->u) { download_rows<double>(uo.p, n, res->u + pb * n, stage); CK(cudaStreamSynchronize(stream)); }
    }
    auto dl = [&](double *host, const double *dev) {
        if (host) CK(cudaMemcpyAsync(host + pb, dev, sizeof(double) * batch, cudaMemcpyDeviceToHost, stream));
    };
    if (res->iters) CK(cudaMemcpyAsync(res->iters + pb, iters.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, stream));
    if (res->status) CK(cudaMemcpyAsync(res->status + pb, status.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, stream));
    dl(res->r_norm, fin.p);
    dl(res->s_norm, fin.p + ld);
    dl(res->eps_pri, fin.p + 2 * ld);
    dl(res->eps_dual, fin.p + 3 * ld);
    dl(res->rho, rho.p);
    CK(cudaStreamSynchronize(stream));
    if (hist_alloc && res->hist_r) {
        const size_t hs = (size_t)max_iter_alloc * ld;
        double *outs[5] = {res->hist_r, res->hist_s, res->hist_eps_pri, res->hist_eps_dual, res->hist_rho};
        for (int a = 0; a < 5; ++a) {
            if (!outs[a]) continue;
            download_rows<double>(hist.p + a * hs, max_iter_alloc, outs[a] + pb * max_iter_alloc, stage);
            CK(cudaStreamSynchronize(stream));
        }
    }
}

}  // namespace

#include "dense_impl.cuh"

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
// The final statistics gather of a multi-GPU handle: one NCCL all-reduce of four 64-bit integers per GPU --
// {converged, sum of iterations, refactorisations} summed, {max iterations} maximised -- over NVLink; the only bytes of
// a solve that cross between GPUs.  libnccl is opened at run time (dlopen: the library has no link-time dependency on
// it); where it is absent the same four integers are summed on the host.
struct NcclApi {
    typedef struct ncclComm *comm_t;
    void *lib = nullptr;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static constexpr int kInt64 = 4, kSum = 0, kMax = 2;      // ncclInt64, ncclSum, ncclMax (nccl.h, stable since NCCL 2.0)
    bool open()
    {
        if (lib) return true;
        if (getenv("ADMMB_NO_NCCL")) return false;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString) return true;
        dlclose(lib);
        lib = nullptr;
        return false;
    }
};

struct admmb_ctx {
    std::vector<Shard> shards;
    NcclApi nccl;
    std::vector<NcclApi::comm_t> comms;      // one communicator per shard (ncclCommInitAll), created at the first gather
    int nccl_state = 0;                      // 0: not tried, 1: communicators ready, -1: unavailable (host sum)
    std::vector<DevBuf<long long>> stat_buf; // per shard: [0..3] send {conv, iters, refac, 0}, [4] send max, [5..8] recv, [9] recv max
    int64_t stat_gathers = 0;                // all-reduces done (tests)
    std::string err;
    std::mutex mu;
    bool uploaded = false;
    int64_t batch = 0;
    int max_iter = 0;
    admmb_opts up_opts;          // the options of the last successful admmb_upload: they fixed the structural choices
                                 // (shared factor or not, dense / tensor-core state, history buffers) that a run must keep
};

namespace {

int fail(admmb_ctx *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

int cuda_code(cudaError_t e)
{
    if (e == cudaErrorMemoryAllocation) return ADMMB_E_NOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return ADMMB_E_NODEVICE;
    return ADMMB_E_CUDA;
}

int validate_opts(admmb_ctx *h, const admmb_opts *op);

int validate(admmb_ctx *h, const admmb_problem *pb, const admmb_opts *op, const admmb_generator *gen = nullptr, bool scp = false)
{
    if (!pb || !op) return fail(h, ADMMB_E_BADARG, "null problem/opts");
    if (gen) {
        if (gen->kind < ADMMB_GEN_CW_IMPULSIVE || gen->kind > ADMMB_GEN_ELLIPTIC_ZOH)
            return fail(h, ADMMB_E_BADARG, "generator kind must be an ADMMB_GEN_* code");
        if (!(gen->T > 0.0) || !std::isfinite(gen->T)) return fail(h, ADMMB_E_BADARG, "generator: stage length T must be > 0");
        if (gen->substeps < 0 || gen->substeps > 4096) return fail(h, ADMMB_E_BADARG, "generator: substeps out of range");
        if (gen->kind == ADMMB_GEN_ELLIPTIC_ZOH) {
            if (!gen->e || !gen->theta0) return fail(h, ADMMB_E_BADARG, "generator: e and theta0 are required for ADMMB_GEN_ELLIPTIC_ZOH");
            for (int64_t i = 0; i < pb->batch; ++i)
                if (!(gen->e[i] >= 0.0 && gen->e[i] < 1.0) || !std::isfinite(gen->theta0[i]) || std::fabs(gen->theta0[i]) > 1.0e3)
                    return fail(h, ADMMB_E_BADARG, "generator: problem %lld needs 0 <= e < 1 and |theta0| <= 1000", (long long)i);
        } else if (!(gen->nmm >= 0.0) || !std::isfinite(gen->nmm)) {
            return fail(h, ADMMB_E_BADARG, "generator: mean motion must be > 0 (0 = 1)");
        }
    }
    if (pb->N < 1 || pb->N > 4096) return fail(h, ADMMB_E_BADARG, "N out of range");
    if (pb->batch < 1 || pb->batch > (int64_t)1 << 30) return fail(h, ADMMB_E_BADARG, "batch out of range");
    if ((!gen && !scp && (!pb->A || !pb->B)) || !pb->s0 || !pb->block_type || !pb->block_par)
        return fail(h, ADMMB_E_BADARG, "A, B, s0, block_type and block_par are required");
    const bool model_batched = scp || (gen ? gen->kind == ADMMB_GEN_ELLIPTIC_ZOH : pb->dyn_batched != 0);
    const int nb = 3 * pb->N + 2;
    int nsplit = 0;
    for (int b = 0; b < nb; ++b) {
        int t = pb->block_type[b];
        if (t < 0 || t > ADMMB_BLK_NONE) return fail(h, ADMMB_E_BADARG, "block_type[%d] = %d is not an ADMMB_BLK_* code", b, t);
        nsplit += t != ADMMB_BLK_NONE;
    }
    if (nsplit == 0) return fail(h, ADMMB_E_BADARG, "no split block: nothing for ADMM to do");
    for (int k = 0; k < pb->N; ++k)
        if (pb->block_type[3 * k + 2] == ADMMB_BLK_NONE && !pb->R)
            return fail(h, ADMMB_E_BADARG, "control block %d is unsplit and R is absent: x-update is singular", k);
    if (op->xupdate == ADMMB_XUPDATE_DENSE) {
        const bool has_P = pb->Q || pb->R;
        if (model_batched) return fail(h, ADMMB_E_BADARG, "dense x-update needs a shared model (dyn_batched = 0)");
        if (has_P && (op->adapt_rho || pb->rho0)) return fail(h, ADMMB_E_BADARG, "dense x-update with P != 0 needs one shared rho");
        if (op->history) return fail(h, ADMMB_E_BADARG, "history is not recorded on the dense path");
        if (op->adapt_rho) return fail(h, ADMMB_E_BADARG, "adaptive rho is not implemented on the dense path (use xupdate = auto / riccati)");
    }
    if (8 * (size_t)nb * sizeof(double) + 4 * (size_t)nb + 4096 > 200 * 1024)
        return fail(h, ADMMB_E_BADARG, "N = %d: the block parameter table (%zu KB) does not fit the shared memory the iteration "
                                       "kernels stage it in (N <= 1000)", pb->N, 8 * (size_t)nb * sizeof(double) / 1024);
    return validate_opts(h, op);
}

// the part of the checks that concerns the options alone (also applied to the options of admmb_run)
int validate_opts(admmb_ctx *h, const admmb_opts *op)
{
    if (!(op->rho > 0.0) || !std::isfinite(op->rho)) return fail(h, ADMMB_E_BADARG, "rho must be > 0");
    if (!(op->alpha > 0.0 && op->alpha < 2.0)) return fail(h, ADMMB_E_BADARG, "alpha must be in (0,2)");
    if (!(op->abstol >= 0.0) || !(op->reltol >= 0.0)) return fail(h, ADMMB_E_BADARG, "tolerances must be >= 0");
    if (op->max_iter < 1) return fail(h, ADMMB_E_BADARG, "max_iter must be >= 1");
    if (op->adapt_rho && (!(op->adapt_mu > 1.0) || !(op->adapt_tau > 1.0) || op->adapt_every < 1))
        return fail(h, ADMMB_E_BADARG, "adaptive rho needs mu > 1, tau > 1, every >= 1");
    if (op->xupdate < 0 || op->xupdate > 2) return fail(h, ADMMB_E_BADARG, "xupdate must be an ADMMB_XUPDATE_* code");
    if (op->precision != ADMMB_PREC_FP64 && op->precision != ADMMB_PREC_TF32) return fail(h, ADMMB_E_BADARG, "bad precision");
    if (op->kernel < ADMMB_KERNEL_AUTO || op->kernel > ADMMB_KERNEL_PINT) return fail(h, ADMMB_E_BADARG, "kernel must be an ADMMB_KERNEL_* code");
    // TF32 + dense: the tensor-core path throughout; TF32 + auto: tensor cores are allowed where they are faster
    // (narrow working sets, eligible problems; otherwise the FP64 Riccati kernel runs)
    if (op->precision == ADMMB_PREC_TF32 && op->xupdate == ADMMB_XUPDATE_RICCATI)
        return fail(h, ADMMB_E_BADARG, "TF32 does not apply to xupdate = riccati (use dense or auto)");
    return ADMMB_OK;
}

template <class Fn>
int guarded(admmb_ctx *h, Fn fn)
{
    try {
        return fn();
    } catch (const CudaFail &f) {
        return fail(h, cuda_code(f.e), "CUDA error %s (%s) at %s", cudaGetErrorName(f.e), cudaGetErrorString(f.e), f.what);
    } catch (const std::bad_alloc &) {
        return fail(h, ADMMB_E_NOMEM, "host allocation failed");
    } catch (...) {
        return fail(h, ADMMB_E_CUDA, "unexpected exception");
    }
}

// run fn(shard_index) on every shard, one host thread per GPU when there are several
template <class Fn>
void for_each_shard(admmb_ctx *h, Fn fn)
{
    const int G = (int)h->shards.size();
    if (G == 1) { fn(0); return; }
    std::vector<std::thread> th;
    std::vector<CudaFail> errs(G, CudaFail{cudaSuccess, ""});
    for (int g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            try { fn(g); } catch (const CudaFail &f) { errs[g] = f; } catch (...) { errs[g] = CudaFail{cudaErrorUnknown, "worker"}; }
        });
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (errs[g].e != cudaSuccess) throw errs[g];
}

// the statistics of a multi-GPU solve, all-reduced over NCCL straight from each GPU's device counters
// ([0] refactorisations, [1] converged, [2] sum of iterations, [3] max iterations -- k_stats).  *done stays false when
// libnccl cannot be opened (the caller then sums the per-GPU results on the host); a failing NCCL call is ADMMB_E_NCCL.
int nccl_gather(admmb_ctx *h, int64_t out[4], bool *done)
{
    *done = false;
    const int G = (int)h->shards.size();
    if (G < 2 || h->nccl_state < 0) return ADMMB_OK;
    NcclApi &N = h->nccl;
    if (h->nccl_state == 0) {
        if (!N.open()) { h->nccl_state = -1; return ADMMB_OK; }
        std::vector<int> devs(G);
        for (int g = 0; g < G; ++g) devs[g] = h->shards[g].device;
        h->comms.assign(G, nullptr);
        const int rc = N.CommInitAll(h->comms.data(), G, devs.data());
        if (rc != 0) { h->nccl_state = -1; h->comms.clear(); return fail(h, ADMMB_E_NCCL, "ncclCommInitAll: %s", N.GetErrorString(rc)); }
        h->stat_buf.resize(G);
        for (int g = 0; g < G; ++g) { CK(cudaSetDevice(h->shards[g].device)); h->stat_buf[g].alloc(5); }
        h->nccl_state = 1;
    }
    int rc = N.GroupStart();
    for (int g = 0; g < G && rc == 0; ++g) {
        Shard &s = h->shards[g];
        CK(cudaSetDevice(s.device));
        rc = N.AllReduce(s.counters.p, h->stat_buf[g].p, 4, NcclApi::kInt64, NcclApi::kSum, h->comms[g], s.stream);
        if (rc == 0) rc = N.AllReduce(s.counters.p + 3, h->stat_buf[g].p + 4, 1, NcclApi::kInt64, NcclApi::kMax, h->comms[g], s.stream);
    }
    const int rc2 = N.GroupEnd();
    if (rc != 0 || rc2 != 0) return fail(h, ADMMB_E_NCCL, "ncclAllReduce: %s", N.GetErrorString(rc != 0 ? rc : rc2));
    long long hv[5];
    CK(cudaSetDevice(h->shards[0].device));
    CK(cudaMemcpyAsync(hv, h->stat_buf[0].p, sizeof(hv), cudaMemcpyDeviceToHost, h->shards[0].stream));
    for (int g = 0; g < G; ++g) { CK(cudaSetDevice(h->shards[g].device)); CK(cudaStreamSynchronize(h->shards[g].stream)); }
    out[0] = hv[1]; out[1] = hv[2]; out[2] = hv[4]; out[3] = hv[0];
    ++h->stat_gathers;
    *done = true;
    return ADMMB_OK;
}

void shard_range(int64_t batch, int G, int g, int64_t &begin, int64_t &cnt)
{
    const int64_t per = (batch + G - 1) / G;
    begin = std::min<int64_t>(batch, per * g);
    cnt = std::min<int64_t>(batch, per * (g + 1)) - begin;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int admmb_version(void) { return ADMMB_VERSION; }

int admmb_create(admmb_handle *out, const int *device_ids, int n_devices)
{
    if (!out) return fail(nullptr, ADMMB_E_BADARG, "null handle pointer");
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible <= 0)
        return fail(nullptr, ADMMB_E_NODEVICE, "no usable CUDA device (%s); libadmm_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (n_devices <= 0) n_devices = device_ids ? 1 : visible;
    admmb_ctx *h = new (std::nothrow) admmb_ctx();
    if (!h) return fail(nullptr, ADMMB_E_NOMEM, "host allocation failed");
    int rc = guarded(nullptr, [&]() {
        h->shards.resize(n_devices);
        for (int g = 0; g < n_devices; ++g) {
            int dev = device_ids ? device_ids[g] : g;
            if (dev < 0 || dev >= visible) { g_create_error = "device id out of range"; return (int)ADMMB_E_NODEVICE; }
            h->shards[g].init(dev);
        }
        return (int)ADMMB_OK;
    });
    if (rc != ADMMB_OK) { delete h; return rc; }
    *out = h;
    return ADMMB_OK;
}

int admmb_destroy(admmb_handle h)
{
    if (!h) return ADMMB_E_BADARG;
    for (auto &s : h->shards) {
        cudaSetDevice(s.device);
        cudaStreamSynchronize(s.stream);
    }
    for (auto c : h->comms)
        if (c && h->nccl.CommDestroy) h->nccl.CommDestroy(c);
    h->comms.clear();
    for (auto &s : h->shards) s.destroy();
    delete h;   // DevBuf destructors release the device memory (cudaFree is device-agnostic under UVA)
    return ADMMB_OK;
}

const char *admmb_last_error(admmb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int admmb_device_count(admmb_handle h) { return h ? (int)h->shards.size() : 0; }

int admmb_nccl_gathers(admmb_handle h) { return h ? (int)h->stat_gathers : 0; }

int admmb_set_stream(admmb_handle h, void *cuda_stream)
{
    if (!h) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    Shard &s = h->shards[0];
    s.stream = cuda_stream ? (cudaStream_t)cuda_stream : s.own_stream;
    return ADMMB_OK;
}

int admmb_upload(admmb_handle h, const admmb_problem *pb, const admmb_opts *op)
{
    return admmb_upload_generated(h, pb, nullptr, op);
}

int admmb_upload_generated(admmb_handle h, const admmb_problem *pb, const admmb_generator *gen, const admmb_opts *op)
{
    if (!h) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->uploaded = false;
    int rc = validate(h, pb, op, gen);
    if (rc != ADMMB_OK) return rc;
    const int G = (int)std::min<int64_t>((int64_t)h->shards.size(), pb->batch);
    rc = guarded(h, [&]() {
        for_each_shard(h, [&](int g) {
            int64_t b, c;
            shard_range(pb->batch, G, g, b, c);
            Shard &s = h->shards[g];
            if (g >= G || c <= 0) { s.uploaded = false; s.batch = 0; return; }
            s.upload(pb, op, b, c, gen);
            if (s.use_dense || s.tf32_tail) dense_prepare(s, op);
        });
        return (int)ADMMB_OK;
    });
    if (rc == ADMMB_OK) { h->uploaded = true; h->batch = pb->batch; h->max_iter = op->max_iter; h->up_opts = *op; }
    return rc;
}

int admmb_run(admmb_handle h, const admmb_opts *op, admmb_result *res)
{
    if (!h || !op) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->uploaded) return fail(h, ADMMB_E_STATE, "admmb_run called before a successful admmb_upload");
    // rho, alpha, the tolerances, chunk, kernel and the adaptation parameters may change between runs of one upload; what
    // upload built its device state on may not (a shared factor computed for ONE rho, dense / tensor-core operands,
    // the size of the history buffers)
    {
        int rc0 = validate_opts(h, op);
        if (rc0 != ADMMB_OK) return rc0;
        const admmb_opts &u0 = h->up_opts;
        if ((op->adapt_rho != 0) != (u0.adapt_rho != 0) || op->xupdate != u0.xupdate || op->precision != u0.precision ||
            (op->history != 0) != (u0.history != 0))
            return fail(h, ADMMB_E_BADARG, "adapt_rho, xupdate, precision and history must be the ones given to admmb_upload "
                                           "(upload the problem again to change them)");
        if (op->history && op->max_iter != u0.max_iter)
            return fail(h, ADMMB_E_BADARG, "max_iter differs from the uploaded one while history is on");
        const Shard &s0 = h->shards[0];
        if ((s0.use_dense || s0.tf32_tail) && (s0.has_Q || s0.has_R) && op->rho != u0.rho)
            return fail(h, ADMMB_E_BADARG, "the dense factor was built for the rho given to admmb_upload (P != 0): upload again to change rho");
    }
    std::vector<admmb_result> part(h->shards.size());
    int rc = guarded(h, [&]() {
        for_each_shard(h, [&](int g) {
            Shard &s = h->shards[g];
            memset(&part[g], 0, sizeof(admmb_result));
            if (s.batch > 0 && s.uploaded) s.run(op, &part[g]);
            else if (h->shards.size() > 1) {       // a GPU without problems still takes part in the statistics all-reduce
                CK(cudaSetDevice(s.device));
                s.counters.alloc(4);
                CK(cudaMemsetAsync(s.counters.p, 0, sizeof(unsigned long long) * 4, s.stream));
            }
        });
        return (int)ADMMB_OK;
    });
    if (rc != ADMMB_OK) return rc;
    if (res) {
        // the final statistics gather: 4 integers per GPU, summed (max for the third) on the host.
        // In the one-process-per-GPU deployment (bench.py under torchrun) this is the NCCL all-reduce.
        res->stats[0] = res->stats[1] = res->stats[2] = res->stats[3] = 0;
        res->device_ms = 0.0;
        res->launches = 0;
        res->kernel_ms = 0.0;
        res->kernel_launches = 0;
        for (size_t g = 0; g < part.size(); ++g) {
            res->kernel_ms = std::max(res->kernel_ms, part[g].kernel_ms);
            res->kernel_launches += part[g].kernel_launches;
            res->stats[0] += part[g].stats[0];
            res->stats[1] += part[g].stats[1];
            res->stats[2] = std::max(res->stats[2], part[g].stats[2]);
            res->stats[3] += part[g].stats[3];
            res->device_ms = std::max(res->device_ms, part[g].device_ms);
            res->launches += h->shards[g].launches;
        }
        // several GPUs in this process: the same totals by one NCCL all-reduce over NVLink (4 + 1 integers per GPU), read
        // back from GPU 0; they must agree with the host-side sums above
        int64_t tot[4];
        bool done = false;
        int rcn = guarded(h, [&]() { return nccl_gather(h, tot, &done); });
        if (rcn != ADMMB_OK) return rcn;
        if (done) {
            for (int i = 0; i < 4; ++i)
                if (tot[i] != res->stats[i])
                    return fail(h, ADMMB_E_NCCL, "statistics all-reduce disagrees with the per-GPU sums (field %d: %lld vs %lld)", i,
                                (long long)tot[i], (long long)res->stats[i]);
        }
    }
    return ADMMB_OK;
}

int admmb_shift_resolve(admmb_handle h, int32_t k, const double *s0_new, const admmb_opts *op, admmb_result *res)
{
    if (!h || !op) return ADMMB_E_BADARG;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (!h->uploaded) return fail(h, ADMMB_E_STATE, "admmb_shift_resolve called before admmb_upload");
        for (auto &s : h->shards) {
            if (s.batch > 0 && !s.ran) return fail(h, ADMMB_E_STATE, "admmb_shift_resolve called before a solve (admmb_run)");
            if (s.batch > 0 && s.use_dense) return fail(h, ADMMB_E_BADARG, "admmb_shift_resolve needs the Riccati path (xupdate = auto / riccati)");
            if (s.batch > 0 && (k < 1 || k > s.N)) return fail(h, ADMMB_E_BADARG, "shift k = %d outside 1 .. N", (int)k);
        }
        int rc = guarded(h, [&]() {
            for_each_shard(h, [&](int g) {
                Shard &s = h->shards[g];
                if (s.batch > 0 && s.uploaded) s.shift_warm_start(k, s0_new);
            });
            return (int)ADMMB_OK;
        });
        if (rc != ADMMB_OK) return rc;
    }
    return admmb_run(h, op, res);
}

int admmb_scp_solve(admmb_handle h, const admmb_problem *pb, const admmb_scp *sc, const admmb_opts *op, admmb_result *res,
                    admmb_scp_result *out)
{
    if (!h || !res) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->uploaded = false;
    int rc = validate(h, pb, op, nullptr, true);
    if (rc != ADMMB_OK) return rc;
    if (!sc) return fail(h, ADMMB_E_BADARG, "null scp");
    if (sc->model != ADMMB_SCP_NL_CIRCULAR && sc->model != ADMMB_SCP_NL_ELLIPTIC)
        return fail(h, ADMMB_E_BADARG, "scp: model must be an ADMMB_SCP_* code");
    if (sc->model == ADMMB_SCP_NL_ELLIPTIC) {
        if (!sc->e || !sc->theta0) return fail(h, ADMMB_E_BADARG, "scp: e and theta0 are required for ADMMB_SCP_NL_ELLIPTIC");
        if (sc->nmm != 0.0 && sc->nmm != 1.0) return fail(h, ADMMB_E_BADARG, "scp: the elliptic model uses the time unit 1 / mean motion (nmm = 1)");
        for (int64_t i = 0; i < pb->batch; ++i)
            if (!(sc->e[i] >= 0.0 && sc->e[i] < 1.0) || !std::isfinite(sc->theta0[i]) || std::fabs(sc->theta0[i]) > 1.0e3)
                return fail(h, ADMMB_E_BADARG, "scp: problem %lld needs 0 <= e < 1 and |theta0| <= 1000", (long long)i);
    }
    if (!(sc->T > 0.0) || !std::isfinite(sc->T)) return fail(h, ADMMB_E_BADARG, "scp: stage length T must be > 0");
    if (!(sc->R0 > 0.0) || !std::isfinite(sc->R0)) return fail(h, ADMMB_E_BADARG, "scp: orbit radius R0 must be > 0");
    if (!(sc->nmm >= 0.0) || !std::isfinite(sc->nmm)) return fail(h, ADMMB_E_BADARG, "scp: mean motion must be > 0 (0 = 1)");
    if (sc->substeps < 0 || sc->substeps > 4096) return fail(h, ADMMB_E_BADARG, "scp: substeps out of range");
    if (sc->max_pass < 1 || sc->max_pass > 1000) return fail(h, ADMMB_E_BADARG, "scp: max_pass must be in 1 .. 1000");
    if (!(sc->tol_abs >= 0.0) || !(sc->tol_rel >= 0.0)) return fail(h, ADMMB_E_BADARG, "scp: tolerances must be >= 0");
    if (sc->control != ADMMB_SCP_CTRL_ZOH && sc->control != ADMMB_SCP_CTRL_IMPULSIVE)
        return fail(h, ADMMB_E_BADARG, "scp: control must be an ADMMB_SCP_CTRL_* code");
    if (op->adapt_rho || op->history || pb->rho0)
        return fail(h, ADMMB_E_BADARG, "scp: adaptive / per-problem rho and the per-iteration history are not carried across passes");
    if (op->xupdate == ADMMB_XUPDATE_DENSE || op->precision != ADMMB_PREC_FP64)
        return fail(h, ADMMB_E_BADARG, "scp: per-problem linearisations need the FP64 Riccati path (xupdate = auto / riccati)");
    for (auto &s : h->shards) s.launches = 0;
    const int G = (int)std::min<int64_t>((int64_t)h->shards.size(), pb->batch);
    std::vector<admmb_result> part(h->shards.size());
    auto t0 = std::chrono::steady_clock::now();
    std::chrono::steady_clock::time_point t1, t2;
    rc = guarded(h, [&]() {
        for_each_shard(h, [&](int g) {
            int64_t b, c;
            shard_range(pb->batch, G, g, b, c);
            Shard &s = h->shards[g];
            memset(&part[g], 0, sizeof(admmb_result));
            if (g >= G || c <= 0) { s.uploaded = false; s.batch = 0; return; }
            struct Flag { bool &f; ~Flag() { f = false; } } flag{s.scp_upload};
            s.scp_upload = true;
            s.upload(pb, op, b, c, nullptr);
        });
        t1 = std::chrono::steady_clock::now();
        for_each_shard(h, [&](int g) {
            Shard &s = h->shards[g];
            if (s.batch > 0 && s.uploaded) s.scp_solve(sc, op, &part[g]);
        });
        t2 = std::chrono::steady_clock::now();
        admmb_scp_result tmp;
        memset(&tmp, 0, sizeof(tmp));
        admmb_scp_result *o = out ? out : &tmp;
        o->stats[0] = o->stats[1] = o->stats[2] = o->stats[3] = 0;
        o->linearise_ms = 0.0;
        // the shards write disjoint ranges of the caller's buffers; the small statistics are summed here
        for (auto &s : h->shards) {
            if (!(s.batch > 0 && s.uploaded)) continue;
            s.download(res);
            s.scp_download(o);
            o->linearise_ms = std::max(o->linearise_ms, s.scp_lin_ms);
        }
        res->stats[0] = res->stats[1] = res->stats[2] = res->stats[3] = 0;
        res->device_ms = res->kernel_ms = 0.0;
        res->launches = res->kernel_launches = 0;
        for (size_t g = 0; g < part.size(); ++g) {
            res->device_ms = std::max(res->device_ms, part[g].device_ms);
            res->kernel_ms = std::max(res->kernel_ms, part[g].kernel_ms);
            res->kernel_launches += part[g].kernel_launches;
            res->launches += h->shards[g].launches;
        }
        res->stats[0] = o->stats[0];
        res->stats[1] = o->stats[1];
        res->stats[2] = o->stats[2];
        res->stats[3] = 0;
        return (int)ADMMB_OK;
    });
    auto t3 = std::chrono::steady_clock::now();
    if (rc != ADMMB_OK) return rc;
    res->h2d_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    res->d2h_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
    h->uploaded = true;
    h->batch = pb->batch;
    h->max_iter = op->max_iter;
    h->up_opts = *op;
    return ADMMB_OK;
}

int admmb_download(admmb_handle h, admmb_result *res)
{
    if (!h || !res) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->uploaded) return fail(h, ADMMB_E_STATE, "admmb_download called before admmb_upload");
    for (auto &s : h->shards)
        if (s.batch > 0 && !s.ran) return fail(h, ADMMB_E_STATE, "admmb_download called before admmb_run");
    return guarded(h, [&]() {
        for_each_shard(h, [&](int g) {
            Shard &s = h->shards[g];
            if (s.batch > 0 && s.uploaded) s.download(res);
        });
        res->launches = 0;
        for (auto &s : h->shards) res->launches += s.launches;
        return (int)ADMMB_OK;
    });
}

int admmb_solve(admmb_handle h, const admmb_problem *pb, const admmb_opts *op, admmb_result *res)
{
    return admmb_solve_generated(h, pb, nullptr, op, res);
}

int admmb_solve_generated(admmb_handle h, const admmb_problem *pb, const admmb_generator *gen, const admmb_opts *op,
                          admmb_result *res)
{
    if (!h || !res) return ADMMB_E_BADARG;
    for (auto &s : h->shards) s.launches = 0;
    auto t0 = std::chrono::steady_clock::now();
    int rc = admmb_upload_generated(h, pb, gen, op);
    if (rc != ADMMB_OK) return rc;
    auto t1 = std::chrono::steady_clock::now();
    rc = admmb_run(h, op, res);
    if (rc != ADMMB_OK) return rc;
    auto t2 = std::chrono::steady_clock::now();
    rc = admmb_download(h, res);
    auto t3 = std::chrono::steady_clock::now();
    res->h2d_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    res->d2h_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
    return rc;
}

}  // extern "C"

#include "unit_api.cuh"
