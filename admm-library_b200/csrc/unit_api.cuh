// unit_api.cuh -- one-kernel entry points with host buffers in/out (SURVEY 4.2 tier T2): they let
// tests/ compare each kernel with the oracle in isolation.  Included by admm_b200.cu.
#pragma once

namespace {

struct UnitCtx {
    Shard &s;
    explicit UnitCtx(admmb_ctx *h) : s(h->shards[0]) { CK(cudaSetDevice(s.device)); }
    void up(DevBuf<double> &b, const double *host, size_t cnt)
    {
        b.alloc(cnt);
        CK(cudaMemcpyAsync(b.p, host, sizeof(double) * cnt, cudaMemcpyHostToDevice, s.stream));
    }
    // host [batch][R] -> device [R][ld]
    void up_rows(DevBuf<double> &b, DevBuf<double> &stg, const double *host, int64_t batch, int R, size_t ld)
    {
        b.alloc((size_t)R * ld);
        stg.alloc((size_t)batch * R);
        CK(cudaMemsetAsync(b.p, 0, sizeof(double) * R * ld, s.stream));
        CK(cudaMemcpyAsync(stg.p, host, sizeof(double) * (size_t)batch * R, cudaMemcpyHostToDevice, s.stream));
        dim3 grid((unsigned)((batch + 31) / 32), (unsigned)((R + 31) / 32)), block(32, 8);
        k_transpose_in<<<grid, block, 0, s.stream>>>(stg.p, batch, R, b.p, ld, nullptr);
        CK(cudaGetLastError());
    }
    void down_rows(const double *dev, DevBuf<double> &stg, double *host, int64_t batch, int R, size_t ld)
    {
        stg.alloc((size_t)batch * R);
        dim3 grid((unsigned)((batch + 31) / 32), (unsigned)((R + 31) / 32)), block(32, 8);
        k_transpose_out<double><<<grid, block, 0, s.stream>>>(dev, ld, R, batch, stg.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host, stg.p, sizeof(double) * (size_t)batch * R, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    void bdesc(DevBuf<int> &b, const int32_t *bt, int nb)
    {
        std::vector<int> h(nb);
        int slot = 0;
        for (int i = 0; i < nb; ++i) h[i] = bt[i] != BLK_NONE ? (bt[i] | (slot++ << 8)) : bt[i];
        b.alloc(nb);
        CK(cudaMemcpy(b.p, h.data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
    }
};

}  // namespace

extern "C" {

int admmb_k_riccati_factor(admmb_handle h, int32_t N, const double *A, const double *B, const double *c,
                           const double *Q, const double *R, double rho, const int32_t *block_type,
                           double *fac_out)
{
    if (!h || N < 1 || !A || !B || !block_type || !fac_out) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    return guarded(h, [&]() {
        UnitCtx U(h);
        DevBuf<double> dA, dB, dc, dQ, dR, dfac;
        DevBuf<int> bd;
        U.up(dA, A, 36 * (size_t)N);
        U.up(dB, B, 18 * (size_t)N);
        if (c) U.up(dc, c, 6 * (size_t)N);
        if (Q) U.up(dQ, Q, 36 * (size_t)(N + 1));
        if (R) U.up(dR, R, 9 * (size_t)N);
        U.bdesc(bd, block_type, 3 * N + 2);
        dfac.alloc((size_t)FS * N);
        k_riccati_factor<<<1, 32, 0, U.s.stream>>>(N, 1, 0, 0, dA.p, dB.p, dc.p, dQ.p, dR.p, 32, nullptr, rho, bd.p,
                                                    dfac.p, nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(fac_out, dfac.p, sizeof(double) * FS * N, cudaMemcpyDeviceToHost, U.s.stream));
        CK(cudaStreamSynchronize(U.s.stream));
        return (int)ADMMB_OK;
    });
}

int admmb_k_xupdate_riccati(admmb_handle h, int32_t N, int64_t batch, const double *fac, int32_t has_c,
                            const double *s0, const double *rt, double *x)
{
    if (!h || N < 1 || batch < 1 || !fac || !s0 || !rt || !x) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    return guarded(h, [&]() {
        UnitCtx U(h);
        const int n = 9 * N + 6;
        const size_t ld = round_up((size_t)batch, 32);
        DevBuf<double> dfac, ds0, drt, dd, dx, stg;
        U.up(dfac, fac, (size_t)FS * N);
        U.up_rows(ds0, stg, s0, batch, 6, ld);
        U.up_rows(drt, stg, rt, batch, n, ld);
        dd.alloc((size_t)3 * N * ld);
        dx.alloc((size_t)n * ld);
        const unsigned gb = (unsigned)((batch + 127) / 128);
        if (has_c) k_xupdate_riccati<true><<<gb, 128, 0, U.s.stream>>>(N, batch, ld, dfac.p, ds0.p, drt.p, dd.p, dx.p);
        else k_xupdate_riccati<false><<<gb, 128, 0, U.s.stream>>>(N, batch, ld, dfac.p, ds0.p, drt.p, dd.p, dx.p);
        CK(cudaGetLastError());
        U.down_rows(dx.p, stg, x, batch, n, ld);
        return (int)ADMMB_OK;
    });
}

int admmb_k_prox_dual_residuals(admmb_handle h, int32_t N, int64_t batch, const int32_t *block_type,
                                const double *block_par, int32_t par_batched, const double *rinv,
                                double alpha, const double *x, double *z, double *u, double *norms)
{
    if (!h || N < 1 || batch < 1 || !block_type || !block_par || !rinv || !x || !z || !u || !norms) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    return guarded(h, [&]() {
        UnitCtx U(h);
        const int n = 9 * N + 6, nb = 3 * N + 2;
        const size_t ld = round_up((size_t)batch, 32);
        DevBuf<double> dpar, drinv, dx, dz, du, dn, stg;
        DevBuf<int> bd;
        U.bdesc(bd, block_type, nb);
        if (par_batched) U.up_rows(dpar, stg, block_par, batch, 8 * nb, ld);
        else U.up(dpar, block_par, (size_t)8 * nb);
        drinv.alloc(ld);
        CK(cudaMemsetAsync(drinv.p, 0, sizeof(double) * ld, U.s.stream));
        CK(cudaMemcpyAsync(drinv.p, rinv, sizeof(double) * batch, cudaMemcpyHostToDevice, U.s.stream));
        U.up_rows(dx, stg, x, batch, n, ld);
        U.up_rows(dz, stg, z, batch, n, ld);
        U.up_rows(du, stg, u, batch, n, ld);
        dn.alloc(5 * ld);
        DenseStep ds;
        memset(&ds, 0, sizeof(ds));
        k_prox_dual_residuals<false, false><<<(unsigned)((batch + 127) / 128), 128, 0, U.s.stream>>>(
            nb, batch, ld, bd.p, dpar.p, par_batched, drinv.p, 0.0, alpha, dx.p, dz.p, du.p, dn.p, nullptr, nullptr, nullptr,
            nullptr, 0, ds);
        CK(cudaGetLastError());
        U.down_rows(dz.p, stg, z, batch, n, ld);
        U.down_rows(du.p, stg, u, batch, n, ld);
        U.down_rows(dn.p, stg, norms, batch, 5, ld);
        return (int)ADMMB_OK;
    });
}

int admmb_k_dense_factor(admmb_handle h, int32_t N, const double *fac, int32_t has_c, double *M, double *S,
                         double *mc)
{
    if (!h || N < 1 || !fac || !M || !S || !mc) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    return guarded(h, [&]() {
        UnitCtx U(h);
        const int n = 9 * N + 6;
        Shard &s = U.s;
        const int keepN = s.N, keepn = s.n;
        s.N = N; s.n = n;
        DevBuf<double> dfac, dM, dS, dmc;
        U.up(dfac, fac, (size_t)FS * N);
        dM.alloc((size_t)n * n); dS.alloc((size_t)n * 6); dmc.alloc(n);
        dense_build_factor(s, dfac.p, has_c != 0, dM.p, dS.p, dmc.p);
        s.N = keepN; s.n = keepn;
        CK(cudaMemcpy(M, dM.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(S, dS.p, sizeof(double) * n * 6, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(mc, dmc.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
        return (int)ADMMB_OK;
    });
}

int admmb_k_xupdate_dense(admmb_handle h, int32_t N, int64_t batch, const double *M, const double *S,
                          const double *mc, const double *s0, const double *rt, int32_t precision, double *x)
{
    if (!h || N < 1 || batch < 1 || !M || !S || !mc || !s0 || !rt || !x) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    return guarded(h, [&]() {
        UnitCtx U(h);
        const int n = 9 * N + 6;
        const size_t ld = round_up((size_t)batch, 32);
        DevBuf<double> dM, dS, dmc, ds0, drt, dx, stg;
        U.up(dM, M, (size_t)n * n);
        U.up(dS, S, (size_t)n * 6);
        U.up(dmc, mc, n);
        U.up_rows(ds0, stg, s0, batch, 6, ld);
        U.up_rows(drt, stg, rt, batch, n, ld);
        dx.alloc((size_t)n * ld);
        if (precision == ADMMB_PREC_TF32 || precision == 2) {
            // precision 1: 3xTF32 split (what the solver uses); 2: plain single-pass TF32
            int rc = dense_tf32_unit(U.s, n, batch, ld, dM.p, dS.p, dmc.p, ds0.p, drt.p, dx.p, precision == 2 ? 1 : 3);
            if (rc != ADMMB_OK) return fail(h, rc, "TF32 dense x-update failed");
        } else {
            dim3 gg((unsigned)((batch + DG_BN - 1) / DG_BN), (unsigned)((n + DG_BM - 1) / DG_BM));
            k_dense_xupdate_f64<<<gg, 256, 0, U.s.stream>>>(n, batch, ld, dM.p, dS.p, dmc.p, ds0.p, drt.p, nullptr, dx.p);
            CK(cudaGetLastError());
        }
        U.down_rows(dx.p, stg, x, batch, n, ld);
        return (int)ADMMB_OK;
    });
}

int admmb_k_generate(admmb_handle h, int32_t N, int64_t batch, const admmb_generator *gen, double *A, double *B)
{
    if (!h || N < 1 || batch < 1 || !gen || !A || !B) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    {
        if (gen->kind < ADMMB_GEN_CW_IMPULSIVE || gen->kind > ADMMB_GEN_ELLIPTIC_ZOH || !(gen->T > 0.0) || gen->substeps < 0 ||
            (gen->kind == ADMMB_GEN_ELLIPTIC_ZOH && (!gen->e || !gen->theta0)))
            return fail(h, ADMMB_E_BADARG, "admmb_k_generate: bad generator");
    }
    return guarded(h, [&]() {
        UnitCtx U(h);
        const bool pp = gen->kind == ADMMB_GEN_ELLIPTIC_ZOH;
        const size_t ld = pp ? round_up((size_t)batch, 32) : 1;
        DevBuf<double> dA, dB, de, dth, stg;
        dA.alloc((size_t)36 * N * ld);
        dB.alloc((size_t)18 * N * ld);
        U.s.generate_model(gen, 0, batch, N, ld, dA.p, dB.p, de, dth);
        if (pp) {
            U.down_rows(dA.p, stg, A, batch, 36 * N, ld);
            U.down_rows(dB.p, stg, B, batch, 18 * N, ld);
        } else {
            CK(cudaMemcpyAsync(A, dA.p, sizeof(double) * 36 * N, cudaMemcpyDeviceToHost, U.s.stream));
            CK(cudaMemcpyAsync(B, dB.p, sizeof(double) * 18 * N, cudaMemcpyDeviceToHost, U.s.stream));
            CK(cudaStreamSynchronize(U.s.stream));
        }
        return (int)ADMMB_OK;
    });
}

int admmb_k_scp_linearise(admmb_handle h, int32_t N, int64_t batch, const admmb_scp *sc, int32_t shoot, const double *s0,
                          double *xref, double *A, double *B, double *c)
{
    if (!h || N < 1 || batch < 1 || !sc || !xref || !A || !B || !c || (shoot && !s0)) return ADMMB_E_BADARG;
    std::lock_guard<std::mutex> lk(h->mu);
    if ((sc->model != ADMMB_SCP_NL_CIRCULAR && sc->model != ADMMB_SCP_NL_ELLIPTIC) || !(sc->T > 0.0) || !(sc->R0 > 0.0) ||
        sc->substeps < 0 || (sc->model == ADMMB_SCP_NL_ELLIPTIC && (!sc->e || !sc->theta0)))
        return fail(h, ADMMB_E_BADARG, "admmb_k_scp_linearise: bad scp parameters");
    return guarded(h, [&]() {
        UnitCtx U(h);
        const size_t ld = round_up((size_t)batch, 32);
        const int n = 9 * N + 6;
        DevBuf<double> dA, dB, dc, ds0, dx, stg;
        dA.alloc((size_t)36 * N * ld);
        dB.alloc((size_t)18 * N * ld);
        dc.alloc((size_t)6 * N * ld);
        U.up_rows(dx, stg, xref, batch, n, ld);
        if (shoot) U.up_rows(ds0, stg, s0, batch, 6, ld);
        ScpConst C;
        const double nmm = sc->nmm != 0.0 ? sc->nmm : 1.0;
        C.substeps = sc->substeps > 0 ? sc->substeps : 8;
        C.R0 = sc->R0; C.twoR0 = 2.0 * sc->R0; C.R0sq = sc->R0 * sc->R0;
        C.n2 = nmm * nmm; C.tn = 2.0 * nmm;
        C.dt = sc->T / (double)C.substeps; C.hdt = 0.5 * C.dt; C.dt6 = C.dt / 6.0;
        C.impulsive = sc->control == ADMMB_SCP_CTRL_IMPULSIVE;
        C.elliptic = sc->model == ADMMB_SCP_NL_ELLIPTIC;
        const unsigned gb = (unsigned)((batch + 127) / 128);
        DevBuf<double> de, dth0, dtab;
        if (C.elliptic) {
            U.up(de, sc->e, (size_t)batch);
            U.up(dth0, sc->theta0, (size_t)batch);
            dtab.alloc((size_t)N * ld);
            k_scp_theta<<<gb, 128, 0, U.s.stream>>>(batch, N, ld, C.dt, C.hdt, C.dt6, C.substeps, de.p, dth0.p, dtab.p);
        }
        if (shoot) k_scp_shoot<<<gb, 128, 0, U.s.stream>>>(C, batch, N, ld, ds0.p, dx.p, dA.p, dB.p, dc.p, de.p, dtab.p);
        else k_scp_linearise<<<dim3(gb, (unsigned)N), 128, 0, U.s.stream>>>(C, batch, N, ld, nullptr, dx.p, dA.p, dB.p, dc.p, de.p, dtab.p);
        CK(cudaGetLastError());
        U.down_rows(dA.p, stg, A, batch, 36 * N, ld);
        U.down_rows(dB.p, stg, B, batch, 18 * N, ld);
        U.down_rows(dc.p, stg, c, batch, 6 * N, ld);
        if (shoot) U.down_rows(dx.p, stg, xref, batch, n, ld);
        return (int)ADMMB_OK;
    });
}

}  // extern "C"
