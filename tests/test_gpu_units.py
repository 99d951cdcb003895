"""GPU unit tests (-m gpu; SURVEY.md 4.2 tier T2): each kernel alone, through its C-ABI entry point,
against the same function of the canonical-order C oracle.  Bit-exact unless stated."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import assert_bit_identical

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))

# the oracle's own factor-record layout (oracle/admm_ocp_cpu.c): name -> (offset, rows, cols, row stride)
ORACLE_FAC = dict(K=(0, 3, 6, 6), Acl=(18, 6, 6, 6), Hinv=(54, 3, 3, 3), E=(63, 3, 6, 6), A=(81, 6, 6, 6),
                  B=(117, 6, 3, 3), c=(135, 1, 6, 6), chat=(141, 1, 6, 6))


def oracle_factor(cpu, pkg, prob, rho):
    m = cpu.to_matlab_layout(prob)
    N = m["N"]
    fac = np.zeros((N, 148))
    dp = lambda a: None if a is None else a.ctypes.data_as(cpu.c_dp)  # noqa: E731
    L = cpu.lib()
    L.ocp_riccati_factor.argtypes = [C.c_int, cpu.c_dp, cpu.c_dp, cpu.c_dp, cpu.c_dp, cpu.c_dp, C.c_double,
                                     cpu.c_ip, cpu.c_dp]
    L.ocp_riccati_factor(N, dp(m["A"]), dp(m["B"]), dp(m["c"]), dp(m["Q"]), dp(m["R"]), float(rho),
                         m["block_type"].ctypes.data_as(cpu.c_ip), dp(fac))
    return fac, pkg.solver.unpack_factor(fac, ORACLE_FAC)


@pytest.mark.parametrize("case", ["cw", "lqr"])
def test_riccati_factor_kernel(solver, cpu_oracle, pkg, P, case):
    if case == "cw":
        prob, _ = P.cfg2_cw_batch(batch=1, N=50)
        rho = 1.0
    else:
        prob, _ = P.lqr_tracking(batch=1, N=30, seed=3)
        rho = 0.37
    got = pkg.solver.unpack_factor(solver.k_riccati_factor(prob, rho))
    _, ref = oracle_factor(cpu_oracle, pkg, prob, rho)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), f"factor part {k} differs: {np.abs(got[k] - ref[k]).max():.3e}"


@pytest.mark.parametrize("has_c", [False, True])
def test_xupdate_riccati_kernel(solver, cpu_oracle, pkg, P, has_c):
    prob, _ = P.lqr_tracking(batch=1, N=25, seed=4, with_affine=has_c)
    N, n, B = 25, 9 * 25 + 6, 70
    fac_o, parts = oracle_factor(cpu_oracle, pkg, prob, 0.8)
    fac_g = pkg.solver.pack_factor(parts, N)
    rng = np.random.default_rng(5)
    s0, rt = rng.standard_normal((B, 6)), rng.standard_normal((B, n))
    x = solver.k_xupdate_riccati(N, fac_g, has_c, s0, rt)
    L = cpu_oracle.lib()
    L.ocp_xupdate_riccati.argtypes = [C.c_int, cpu_oracle.c_dp, C.c_int, cpu_oracle.c_dp, cpu_oracle.c_dp, cpu_oracle.c_dp]
    ref = np.zeros_like(rt)
    for p in range(B):
        xo = np.zeros(n)
        L.ocp_xupdate_riccati(N, fac_o.ctypes.data_as(cpu_oracle.c_dp), int(has_c),
                              np.ascontiguousarray(s0[p]).ctypes.data_as(cpu_oracle.c_dp),
                              np.ascontiguousarray(rt[p]).ctypes.data_as(cpu_oracle.c_dp), xo.ctypes.data_as(cpu_oracle.c_dp))
        ref[p] = xo
    assert np.array_equal(x, ref)


@pytest.mark.parametrize("par_batched", [False, True])
def test_prox_dual_residual_kernel_all_block_types(solver, cpu_oracle, P, par_batched):
    N, B = 10, 45
    n, nb = 9 * N + 6, 3 * N + 2
    rng = np.random.default_rng(6)
    bt = (np.arange(nb) % 9).astype(np.int32)
    bp = np.zeros((B if par_batched else 1, nb, 8))
    bp[..., 0] = rng.uniform(0.05, 0.5, bp.shape[:2])
    bp[..., 1] = rng.uniform(0.2, 1.0, bp.shape[:2])
    bp[..., 2:5] = rng.uniform(-0.8, -0.1, bp.shape[:2] + (3,))
    bp[..., 5:8] = rng.uniform(0.1, 0.8, bp.shape[:2] + (3,))
    x, z, u = (rng.standard_normal((B, n)) for _ in range(3))
    rinv = rng.uniform(0.5, 2.0, B)
    zg, ug, ng = solver.k_prox_dual_residuals(N, bt, bp, rinv, 1.4, x, z, u)
    L = cpu_oracle.lib()
    dp = cpu_oracle.c_dp
    L.ocp_prox_dual_residuals.argtypes = [C.c_int, cpu_oracle.c_ip, dp, C.c_double, C.c_double, dp, dp, dp, dp]
    for p in range(B):
        zo, uo, no = z[p].copy(), u[p].copy(), np.zeros(5)
        par = np.ascontiguousarray(bp[p if par_batched else 0])
        L.ocp_prox_dual_residuals(nb, bt.ctypes.data_as(cpu_oracle.c_ip), par.ctypes.data_as(dp), float(rinv[p]), 1.4,
                                  np.ascontiguousarray(x[p]).ctypes.data_as(dp), zo.ctypes.data_as(dp),
                                  uo.ctypes.data_as(dp), no.ctypes.data_as(dp))
        split = np.repeat(bt != 8, 3)
        assert np.array_equal(zg[p][split], zo[split]) and np.array_equal(ug[p][split], uo[split])
        assert np.array_equal(zg[p][~split], z[p][~split])       # BLK_NONE rows untouched
        assert np.array_equal(ng[p], no)


def test_dense_factor_and_fp64_dense_xupdate_kernels(solver, cpu_oracle, pkg, P):
    prob, _ = P.lqr_tracking(batch=1, N=12, seed=8)
    N, n, B = 12, 9 * 12 + 6, 150
    fac_o, parts = oracle_factor(cpu_oracle, pkg, prob, 1.3)
    M, S, mc = solver.k_dense_factor(N, pkg.solver.pack_factor(parts, N), True)
    L = cpu_oracle.lib()
    dp = cpu_oracle.c_dp
    Mo, So, mo = np.zeros((n, n)), np.zeros((n, 6)), np.zeros(n)
    L.ocp_kkt_dense_factor.argtypes = [C.c_int, dp, C.c_int, dp, dp, dp]
    L.ocp_kkt_dense_factor(N, fac_o.ctypes.data_as(dp), 1, Mo.ctypes.data_as(dp), So.ctypes.data_as(dp), mo.ctypes.data_as(dp))
    assert np.array_equal(M, Mo) and np.array_equal(S, So) and np.array_equal(mc, mo)
    # the dense factor really is KKT^-1 restricted (independent check against numpy.linalg)
    from oracle import admm_ocp as O
    wblk, _ = O.split_weights(prob["block_type"])
    d = O.kkt_dense_factor(prob["A"][0], prob["B"][0], prob["c"][0], prob["Q"][0], prob["R"][0], 1.3, wblk)
    assert np.allclose(M, d["M"], atol=1e-9) and np.allclose(S, d["S"], atol=1e-9) and np.allclose(mc, d["mc"], atol=1e-9)
    rng = np.random.default_rng(9)
    s0, rt = rng.standard_normal((B, 6)), rng.standard_normal((B, n))
    x = solver.k_xupdate_dense(N, M, S, mc, s0, rt, "fp64")
    L.ocp_xupdate_dense.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp]
    for p in range(0, B, 7):
        xo = np.zeros(n)
        L.ocp_xupdate_dense(n, Mo.ctypes.data_as(dp), So.ctypes.data_as(dp), mo.ctypes.data_as(dp),
                            np.ascontiguousarray(s0[p]).ctypes.data_as(dp), np.ascontiguousarray(rt[p]).ctypes.data_as(dp),
                            xo.ctypes.data_as(dp))
        assert np.array_equal(x[p], xo)


@pytest.mark.parametrize("case", ["lqr", "cw"])
def test_dense_fp64_path_end_to_end(solver, cpu_oracle, P, case):
    if case == "lqr":
        prob, opts = P.lqr_tracking(batch=70, N=10, seed=2)
        opts = dict(opts, max_iter=120, xupdate="dense")
    else:
        prob, opts = P.cfg2_cw_batch(batch=130, N=12, seed=2)
        opts = dict(opts, max_iter=250, xupdate="dense", alpha=1.5)
    got = solver.solve(prob, opts)
    ref = cpu_oracle.solve(prob, opts)
    assert_bit_identical(got, ref, f"dense fp64 {case}")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_gpu_reproduces_golden_fixtures(solver, path):
    """Committed NumPy-oracle vectors: same iteration counts, x/z/u within 1e-9 relative (north_star bar)."""
    g = np.load(path)
    prob = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    prob["N"] = prob["A"].shape[1]
    opts = {k[4:]: g[k].item() for k in g.files if k.startswith("opt_")}
    x, z, u, h = solver.solve(prob, opts)
    assert np.array_equal(h["iters"], g["out_iters"])
    assert np.array_equal(h["status"], g["out_status"])
    for a, b in ((x, g["out_x"]), (z, g["out_z"]), (u, g["out_u"])):
        scale = max(1.0, np.abs(b).max())
        assert np.abs(a - b).max() <= 1e-9 * scale
    assert np.allclose(h["hist"]["r_norm"], g["out_hist_r"], rtol=1e-7, atol=1e-12, equal_nan=True)
    assert h["refactor_count"] == int(g["out_refactor"])


def test_staged_api_and_repeatability(solver, cpu_oracle, P):
    prob, opts = P.cfg2_cw_batch(batch=300, N=20, seed=8)
    opts = dict(opts, max_iter=400)
    solver.upload(prob, opts)
    r1 = solver.run(opts)
    a = solver.download(opts)
    r2 = solver.run(opts)                       # run() restarts from the uploaded warm start
    b = solver.download(opts)
    assert r1["stats"] == r2["stats"] and r1["kernel_launches"] > 0
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    ref = cpu_oracle.solve(prob, opts)
    assert_bit_identical(a, ref, "staged")
    assert r1["stats"][:3] == [int(v) for v in ref[3]["stats"][:3]]
    x_only = solver.download(opts, want=("x",))
    assert x_only[1] is None and np.array_equal(x_only[0], a[0])


def test_bad_arguments_are_refused_not_crashed(solver, pkg, P):
    prob, opts = P.cfg2_cw_batch(batch=4, N=5, seed=1)
    E = pkg._lib.E_BADARG
    for bad in (dict(rho=-1.0), dict(alpha=2.5), dict(max_iter=0), dict(adapt_rho=1, adapt_tau=1.0),
                dict(precision="tf32", xupdate="riccati"), dict(xupdate="dense", history=1),
                dict(xupdate="dense", adapt_rho=1)):
        with pytest.raises(pkg.AdmmError) as e:
            solver.solve(prob, dict(opts, **bad))
        assert e.value.code == E
    bt = prob["block_type"].copy()
    bt[2] = 99
    with pytest.raises(pkg.AdmmError):
        solver.solve(dict(prob, block_type=bt), opts)
    bt = prob["block_type"].copy()
    bt[2] = P.BLK_NONE                                       # unsplit control without R: singular x-update
    with pytest.raises(pkg.AdmmError):
        solver.solve(dict(prob, block_type=bt), opts)
    with pytest.raises(pkg.AdmmError) as e:
        pkg.Solver().download(opts)
    assert e.value.code == pkg._lib.E_STATE
    # a NaN input is a per-problem status, never an error of the call
    prob["s0"][1, 0] = np.nan
    x, z, u, h = solver.solve(prob, dict(opts, max_iter=20))
    assert h["status"][1] == 2 and h["status"][0] != 2


def test_full_size_cfg2_properties(solver, cpu_oracle, P):
    """configs[1] at full size (4,096 x N=50): size-independent properties + oracle parity on a slice."""
    from oracle import admm_ocp as O
    prob, opts = P.cfg2_cw_batch(batch=4096, N=50, seed=2)
    opts = dict(opts, max_iter=600)
    x, z, u, h = solver.solve(prob, opts)
    N, n = 50, 456
    G = O.assemble_G(prob["A"][0], prob["B"][0], N)
    hv = np.zeros((4096, 6 * (N + 1)))
    hv[:, :6] = prob["s0"]
    assert np.abs(x @ G.T - hv).max() < 1e-9                 # every x satisfies the dynamics
    zb = z.reshape(4096, -1, 3)
    assert np.all(np.abs(zb[:, 2:3 * N:3]) <= 0.4 + 1e-15)   # z inside the dv box
    assert np.all(zb[:, 3 * N:] == 0.0)                      # terminal point
    sl = slice(1000, 1064)
    sub = dict(prob, s0=prob["s0"][sl])
    xr, zr, ur, hr = cpu_oracle.solve(sub, opts)
    assert np.array_equal(h["iters"][sl], hr["iters"])
    assert np.array_equal(x[sl], xr) and np.array_equal(z[sl], zr) and np.array_equal(u[sl], ur)


# ----------------------------------------------------------------------------- TF32 tensor-core path
def test_tf32_tcgen05_dense_xupdate_kernel(solver, P):
    """Row a2' on the tensor cores (tcgen05.mma kind::tf32 + TMA + TMEM).  Stated tolerance: 3xTF32 split
    operands (what the solver uses) 1e-5 relative to max|x|; single-pass TF32 2e-3 (10-bit mantissa)."""
    from oracle import admm_ocp as O
    prob, _ = P.cfg2_cw_batch(batch=1, N=50)
    N, n = 50, 456
    wblk, w = O.split_weights(prob["block_type"])
    d = O.kkt_dense_factor(prob["A"][0], prob["B"][0], None, None, None, 1.0, wblk)
    rng = np.random.default_rng(3)
    for B in (100, 1000):
        s0, rt = rng.standard_normal((B, 6)), w * rng.standard_normal((B, n))
        ref = rt @ d["M"].T + s0 @ d["S"].T + d["mc"]
        scale = np.abs(ref).max()
        x3 = solver.k_xupdate_dense(N, d["M"], d["S"], d["mc"], s0, rt, "tf32")
        x1 = solver.k_xupdate_dense(N, d["M"], d["S"], d["mc"], s0, rt, "tf32_single")
        assert np.abs(x3 - ref).max() <= 1e-5 * scale
        assert 1e-6 * scale < np.abs(x1 - ref).max() <= 2e-3 * scale      # really TF32, not a silent FP64 path


@pytest.mark.parametrize("case", ["lqr", "cw"])
def test_tf32_dense_path_end_to_end(solver, cpu_oracle, P, case):
    """precision='tf32': a separate, stated precision class (north_star: 1e-4 where TF32 is used).
    Iteration counts are NOT expected to equal FP64; the converged iterates must agree
    with the FP64 oracle to 1e-4 relative."""
    if case == "lqr":
        prob, opts = P.lqr_tracking(batch=200, N=20, seed=2)
        opts = dict(opts, max_iter=400, abstol=1e-5, reltol=1e-5)
    else:
        prob, opts = P.cfg2_cw_batch(batch=300, N=20, seed=2)
        opts = dict(opts, max_iter=8000, abstol=1e-5, reltol=1e-5)
    x, z, u, h = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32"))
    xr, zr, ur, hr = cpu_oracle.solve(prob, opts)
    both = (h["status"] == 0) & (hr["status"] == 0)
    assert both.mean() > 0.7
    sx = np.abs(xr[both]).max()
    assert np.abs(x[both] - xr[both]).max() <= 1e-3 * sx       # both are 1e-5-tolerance solutions of the same QP
    assert np.abs(z[both] - zr[both]).max() <= 1e-3 * sx
    it_ratio = h["iters"][both].astype(float) / hr["iters"][both]
    assert 0.5 < np.median(it_ratio) < 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("max_iter", [6, 7])
def test_tf32_condensed_matches_oracle_iterates(solver, cpu_oracle, P, max_iter):
    """Condensed TF32 path (only the split rows in the per-iteration GEMM, two right-hand-side buffers, full x
    from one final GEMM): after a FIXED small number of iterations -- far from convergence, so x^k and x^{k+1}
    differ visibly -- x, z, u must match the FP64 oracle's iterates of the same iteration to TF32x3 accuracy,
    for an even and an odd count (the final GEMM must pick the right buffer), and must agree with the
    uncondensed TF32 path."""
    import os
    prob, opts = P.cfg2_cw_batch(batch=100, N=20, seed=5)
    opts = dict(opts, max_iter=max_iter, abstol=1e-12, reltol=1e-12)
    xr, zr, ur, hr = cpu_oracle.solve(prob, opts)
    x, z, u, h = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32"))
    os.environ["ADMMB_NO_CONDENSED"] = "1"
    try:
        x2, z2, u2, h2 = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32"))
    finally:
        del os.environ["ADMMB_NO_CONDENSED"]
    assert (h["iters"] == max_iter).all() and (h["status"] == 1).all()
    sx = np.abs(xr).max()
    for a, b in ((x, xr), (z, zr), (u, ur), (x, x2), (z, z2), (u, u2)):
        assert np.abs(a - b).max() <= 2e-5 * sx
    # x^k and x^{k+1} are far apart at this point: a wrong buffer would fail the check above
    opts1 = dict(opts, max_iter=max_iter + 1)
    xn = cpu_oracle.solve(prob, opts1)[0]
    assert np.abs(xn - xr).max() > 1e-3 * sx
    np.testing.assert_allclose(h["r_norm"], hr["r_norm"], rtol=1e-3, atol=1e-6 * sx)


@pytest.mark.gpu
def test_tf32_condensed_early_exit_keeps_final_x(solver, cpu_oracle, P):
    """Problems finish at different iterations; each one's x must come from ITS last iteration: x on the split
    rows must reproduce the reported primal residual against z."""
    prob, opts = P.cfg2_cw_batch(batch=256, N=20, seed=9)
    opts = dict(opts, max_iter=4000, abstol=1e-4, reltol=1e-4, chunk=5)
    x, z, u, h = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32"))
    assert (h["status"] == 0).mean() > 0.9 and len(np.unique(h["iters"])) > 5
    bt = np.asarray(prob["block_type"])
    rows = np.repeat(bt != P.BLK_NONE, 3)
    r = np.linalg.norm((x - z)[:, rows], axis=1)
    np.testing.assert_allclose(r, h["r_norm"], rtol=2e-2, atol=1e-5 * np.abs(x).max())


@pytest.mark.gpu
def test_tf32_incremental_reaches_fp64_tolerance(solver, cpu_oracle, P):
    """The condensed path applies the tensor-core GEMM to the INCREMENT of the right-hand side and accumulates x in
    FP64, so its rounding error scales with the ADMM step and it meets the same 1e-6 tolerance as the FP64 path
    (the absolute TF32 form stalls near |r| ~ 1e-6 |x|): same converged set, nearly the same iteration counts,
    iterates within 1e-5, and a returned trajectory that satisfies the dynamics to FP64 round-off."""
    prob, opts = P.cfg2_cw_batch(batch=512, N=50, seed=11)
    assert opts["abstol"] <= 1e-6 and opts["reltol"] <= 1e-6
    xr, zr, ur, hr = cpu_oracle.solve(prob, opts)
    x, z, u, h = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32"))
    # every problem of the (re-tuned, round 2) workload converges in FP64; the TF32 class must converge on the same set,
    # in nearly the same number of iterations in the median (single problems stop up to ~30 % earlier or later: the
    # stopping test of a slowly converging LP is crossed at a shallow angle)
    assert (hr["status"] == 0).all()
    assert (h["status"] == hr["status"]).all()
    both = (h["status"] == 0) & (hr["status"] == 0)
    ratio = h["iters"][both].astype(float) / hr["iters"][both]
    assert 0.98 < np.median(ratio) < 1.02 and 0.6 < ratio.min() and ratio.max() < 1.6
    sx = np.abs(xr).max()
    # these fuel-optimal (L1) problems are nearly degenerate: points that satisfy the 1e-6 stopping test are much farther
    # than 1e-6 apart along flat directions (the FP64 oracle's own x still moves by ~2e-4 |x| when its tolerance is
    # tightened to 1e-9), so the converged points are compared through the objective and a matching, looser bound on x
    from oracle import admm_ocp as O
    fz, fr = O.objective(prob, z)[both], O.objective(prob, zr)[both]
    assert np.abs(fz - fr).max() <= 1e-4 * np.abs(fr).max()
    assert np.abs(x[both] - xr[both]).max() <= 5e-3 * sx
    # dynamics residual of the returned x: s_{k+1} - A s_k - B a_k
    A, B = np.asarray(prob["A"])[0], np.asarray(prob["B"])[0]
    N = A.shape[0]
    X = x[:, :9 * N].reshape(-1, N, 9)
    s, a = X[:, :, :6], X[:, :, 6:]
    s_next = np.concatenate([s[:, 1:], x[:, 9 * N:].reshape(-1, 1, 6)], axis=1)
    res = s_next - np.einsum("kij,pkj->pki", A, s) - np.einsum("kij,pkj->pki", B, a)
    assert np.abs(res).max() <= 1e-12 * sx


@pytest.mark.gpu
def test_tf32_converged_point_within_1e4_of_fp64(solver, cpu_oracle, P):
    """north_star: "final x/z/u within ... 1e-4 where TF32 is used, stated".  Both sides are solved far beyond the working
    tolerance (1e-9), so that the comparison is between converged POINTS and not between two places on the flat bottom
    of a nearly degenerate LP at which a 1e-6 stopping test happens to fire: equal converged sets and
    max |x - x_ref|, |z - z_ref| <= 1e-4 |x|."""
    prob, opts = P.cfg2_cw_batch(batch=256, N=50, seed=12)
    o = dict(opts, abstol=1e-9, reltol=1e-9, max_iter=150000)
    xr, zr, ur, hr = cpu_oracle.solve(prob, o)
    x, z, u, h = solver.solve(prob, dict(o, xupdate="dense", precision="tf32"))
    assert (hr["status"] == 0).mean() > 0.99
    assert np.array_equal(h["status"], hr["status"])
    both = (h["status"] == 0) & (hr["status"] == 0)
    sx = np.abs(xr).max()
    err_x, err_z = np.abs(x[both] - xr[both]).max() / sx, np.abs(z[both] - zr[both]).max() / sx
    print(f"tf32 vs fp64 at tolerance 1e-9: max |dx| / |x| = {err_x:.2e}, max |dz| / |x| = {err_z:.2e}")
    assert err_x <= 1e-4 and err_z <= 1e-4


@pytest.mark.gpu
def test_tf32_incremental_reaches_1e8_where_absolute_form_stalls(solver, P):
    """Tolerance 1e-8: the incremental form (with its periodic exact refresh of x_R) converges like the FP64 Riccati
    path; the absolute TF32x3 form (ADMMB_NO_CONDENSED=1: X = M RT on the tensor cores) stalls on its ~1e-6 |x|
    rounding floor.  scripts/tf32_tolerance_sweep.py prints the full table."""
    import os
    prob, opts = P.cfg2_cw_batch(batch=128, N=50, seed=11)
    o = dict(opts, abstol=1e-8, reltol=1e-8, max_iter=40000)
    ref = solver.solve(prob, o)[3]
    inc = solver.solve(prob, dict(o, xupdate="dense", precision="tf32"))[3]
    os.environ["ADMMB_NO_CONDENSED"] = "1"
    try:
        ab = solver.solve(prob, dict(o, xupdate="dense", precision="tf32"))[3]
    finally:
        del os.environ["ADMMB_NO_CONDENSED"]
    n_ref, n_inc, n_abs = [(h["status"] == 0).sum() for h in (ref, inc, ab)]
    assert n_ref > 0.6 * 128
    assert abs(int(n_inc) - int(n_ref)) <= 4
    assert n_abs < 0.2 * n_ref
    both = (ref["status"] == 0) & (inc["status"] == 0)
    ratio = inc["iters"][both].astype(float) / ref["iters"][both]
    assert 0.97 < np.median(ratio) < 1.03



@pytest.mark.gpu
@pytest.mark.parametrize("switch", [700, 100000])
def test_tf32_auto_riccati_then_tensor_core_tail(solver, cpu_oracle, P, switch):
    """precision='tf32' with xupdate='auto': the FP64 Riccati kernel runs while the working set is wide, the condensed
    incremental tensor-core pair takes over the compacted working set once it is narrow (switch=700: after ~half of
    the 1,500 problems have finished; switch=100000: from iteration 0).  Problems that finish in the first phase are
    bit-identical to the oracle; the others follow it within the TF32 class; every returned x satisfies the dynamics
    to FP64 round-off (it is rebuilt from the backward-sweep result d by the Riccati path's output kernel)."""
    import os
    prob, opts = P.cfg2_cw_batch(batch=1500, N=20, seed=31)
    opts = dict(opts, max_iter=6000, abstol=1e-6, reltol=1e-6)
    xr, zr, ur, hr = cpu_oracle.solve(prob, opts)
    x, z, u, h = solver.solve(prob, dict(opts, precision="tf32", tf32_switch=switch))
    assert (hr["status"] == 0).mean() > 0.6
    assert (h["status"] == hr["status"]).mean() > 0.99
    both = (h["status"] == 0) & (hr["status"] == 0)
    di = np.abs(h["iters"][both].astype(int) - hr["iters"][both].astype(int))
    assert (di <= 0.02 * hr["iters"][both] + 2).mean() > 0.98
    if switch == 700:
        # the first finishers never saw the tensor cores: bit-identical
        early = both & (hr["iters"] <= np.sort(hr["iters"][both])[200])
        assert early.sum() >= 200
        np.testing.assert_array_equal(h["iters"][early], hr["iters"][early])
        np.testing.assert_array_equal(x[early], xr[early])
        np.testing.assert_array_equal(z[early], zr[early])
        np.testing.assert_array_equal(u[early], ur[early])
    from oracle import admm_ocp as O
    fz, fr = O.objective(prob, z)[both], O.objective(prob, zr)[both]
    assert np.abs(fz - fr).max() <= 1e-4 * np.abs(fr).max()
    sx = np.abs(xr).max()
    assert np.abs(x[both] - xr[both]).max() <= 5e-3 * sx
    A, B = np.asarray(prob["A"])[0], np.asarray(prob["B"])[0]
    N = A.shape[0]
    X = x[:, :9 * N].reshape(-1, N, 9)
    s_, a_ = X[:, :, :6], X[:, :, 6:]
    s_next = np.concatenate([s_[:, 1:], x[:, 9 * N:].reshape(-1, 1, 6)], axis=1)
    res = s_next - np.einsum("kij,pkj->pki", A, s_) - np.einsum("kij,pkj->pki", B, a_)
    assert np.abs(res).max() <= 1e-12 * sx
    # the reported residual (from the accumulated x_R) matches the returned iterates (exact x - z on the split rows)
    # up to the rounding x_R has accumulated since its last refresh
    rows = np.repeat(np.asarray(prob["block_type"]) != P.BLK_NONE, 3)
    r = np.linalg.norm((x - z)[:, rows], axis=1)
    np.testing.assert_allclose(r[both], h["r_norm"][both], rtol=5e-2, atol=3e-8 * sx)



@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["affine", "single_pass", "batched_par"])
def test_tf32_condensed_variants(solver, cpu_oracle, P, variant):
    """Less-travelled branches of the condensed tensor-core path: affine dynamics (c != 0: k_tf32_final_x<true, ...> and
    the x_R initialisation carry the offset, the increments do not), single-pass TF32 operands (ADMMB_TF32_SINGLE: no
    lo buffers), per-problem prox parameters (the parameter array travels with the working set)."""
    import os
    prob, opts = P.cfg2_cw_batch(batch=200, N=20, seed=41)
    opts = dict(opts, max_iter=5000, abstol=1e-5, reltol=1e-5)
    prob = dict(prob)
    if variant == "affine":
        rng = np.random.default_rng(3)
        prob["c"] = 1e-3 * rng.standard_normal((1, 20, 6))
    if variant == "batched_par":
        rng = np.random.default_rng(4)
        bp = np.repeat(np.asarray(prob["block_par"]), 200, axis=0)
        bp[:, :, P.PAR_LAM] *= rng.uniform(0.5, 2.0, size=(200, 1))
        prob["block_par"] = bp
    xr, zr, ur, hr = cpu_oracle.solve(prob, opts)
    if variant == "single_pass":
        os.environ["ADMMB_TF32_SINGLE"] = "1"
    try:
        x, z, u, h = solver.solve(prob, dict(opts, xupdate="dense", precision="tf32", chunk=20))
    finally:
        os.environ.pop("ADMMB_TF32_SINGLE", None)
    assert (hr["status"] == 0).mean() > 0.5
    # (single-pass operands perturb every increment by ~7e-4 of the step: slow problems near max_iter may flip)
    assert (h["status"] == hr["status"]).mean() > (0.85 if variant == "single_pass" else 0.97)
    both = (h["status"] == 0) & (hr["status"] == 0)
    ratio = h["iters"][both].astype(float) / hr["iters"][both]
    assert 0.9 < np.median(ratio) < 1.1
    from oracle import admm_ocp as O
    fz, fr = O.objective(prob, z)[both], O.objective(prob, zr)[both]
    assert np.abs(fz - fr).max() <= (1e-2 if variant == "single_pass" else 1e-3) * np.abs(fr).max()
    sx = np.abs(xr).max()
    assert np.abs(x[both] - xr[both]).max() <= 2e-2 * sx
    # dynamics of the returned x, incl. the affine term
    A, B = np.asarray(prob["A"])[0], np.asarray(prob["B"])[0]
    N = A.shape[0]
    X = x[:, :9 * N].reshape(-1, N, 9)
    s_, a_ = X[:, :, :6], X[:, :, 6:]
    s_next = np.concatenate([s_[:, 1:], x[:, 9 * N:].reshape(-1, 1, 6)], axis=1)
    res = s_next - np.einsum("kij,pkj->pki", A, s_) - np.einsum("kij,pkj->pki", B, a_)
    if prob.get("c") is not None:
        res = res - np.asarray(prob["c"])[0][None]
    assert np.abs(res).max() <= 1e-12 * sx


@pytest.mark.gpu
def test_run_rejects_options_that_contradict_the_upload(pkg, P):
    """admmb_run must not accept options that change what admmb_upload built on the device (a shared factor
    computed for one rho with per-problem adaptive rho, history buffers of another size, another x-update)."""
    prob, opts = P.lqr_tracking(batch=16, N=8, seed=3)            # Q, R present: the factor depends on rho
    with pkg.Solver() as s:
        s.upload(prob, dict(opts, adapt_rho=0, history=0))
        s.run(dict(opts, adapt_rho=0, history=0, rho=2.0, alpha=1.2, max_iter=77))      # free to vary
        for bad in (dict(adapt_rho=1), dict(history=1), dict(xupdate="dense"), dict(rho=-1.0), dict(alpha=2.5)):
            with pytest.raises(pkg._lib.AdmmError) as e:
                s.run({**opts, "adapt_rho": 0, "history": 0, **bad})
            assert e.value.code == pkg._lib.E_BADARG, bad
        s.upload(prob, dict(opts, history=1, max_iter=50))
        with pytest.raises(pkg._lib.AdmmError):
            s.run(dict(opts, history=1, max_iter=60))


@pytest.mark.gpu
def test_history_of_iterations_never_run_is_nan_on_the_c_abi(pkg, P):
    """callers of the C ABI / MEX gateway get the history as the device wrote it; entries of iterations a
    problem never ran must be NaN (as in the oracle), not stale device memory."""
    prob, opts = P.cfg2_cw_batch(batch=40, N=10, seed=9)
    o = dict(opts, rho=1.0, alpha=1.6, max_iter=300, history=1)
    with pkg.Solver() as s:
        for _ in range(2):                                         # second run: the buffer holds an older history
            s.upload(prob, o)
            s.run(o)
            res = pkg.solver.ResultBuffers(40, 96, 300, True)
            for v in res.hist.values():
                v[:] = 0.0                                         # not NaN: whatever is NaN afterwards came from the device
            s.download_c(res.c)
            it = res.iters
            k = np.arange(300)[None, :]
            for name, v in res.hist.items():
                assert np.isnan(v[k >= it[:, None]]).all(), name
                assert np.isfinite(v[k < it[:, None]]).all(), name


# ---- on-device generators ---------------------------------------------------------------------------
GEN_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "generators", "*_*.npz")))


@pytest.mark.parametrize("path", [g for g in GEN_GOLDEN if "sincos" not in g],
                         ids=[os.path.basename(g)[:-4] for g in GEN_GOLDEN if "sincos" not in g])
def test_generator_kernels_reproduce_golden_fixtures_bit_for_bit(solver, path):
    g = np.load(path)
    kind = str(g["kind"])
    gen = dict(kind=kind, T=float(g["T"]))
    if kind == "elliptic_zoh":
        gen.update(e=g["e"], theta0=g["theta0"], substeps=int(g["substeps"]))
        batch = len(g["e"])
    else:
        gen.update(nmm=float(g["nmm"]))
        batch = 1
    A, B = solver.k_generate(int(g["N"]), batch, gen)
    assert np.array_equal(A, g["A"]), f"A differs: {np.abs(A - g['A']).max():.3e}"
    assert np.array_equal(B, g["B"]), f"B differs: {np.abs(B - g['B']).max():.3e}"


@pytest.mark.parametrize("batch,N,sub", [(1, 1, 1), (33, 3, 8), (257, 50, 8), (4096, 20, 4)])
def test_generator_elliptic_matches_oracle_generator(solver, batch, N, sub):
    from oracle import gen_ocp as G
    rng = np.random.Generator(np.random.PCG64(100 + batch))
    e, th = rng.uniform(0.0, 0.75, batch), rng.uniform(-7.0, 7.0, batch)
    T = 2 * np.pi / N
    A, B = solver.k_generate(N, batch, dict(kind="elliptic_zoh", T=T, e=e, theta0=th, substeps=sub))
    Ar, Br = G.elliptic_stage_matrices(e, th, N, T, sub)
    assert np.array_equal(A, Ar) and np.array_equal(B, Br)


@pytest.mark.parametrize("kind", ["elliptic_zoh", "cw_impulsive", "cw_zoh"])
def test_solve_generated_equals_solve_with_uploaded_oracle_matrices(solver, cpu_oracle, P, kind):
    """The generated model feeds the same factor / iteration kernels: the solve is bit-identical to the oracle's solve
    of the problem whose (A, B) come from the oracle generator."""
    from oracle import gen_ocp as G
    if kind == "elliptic_zoh":
        prob, opts = P.cfg4_elliptic(batch=97, N=14, seed=41)
        gen = dict(kind=kind, T=2 * np.pi / 14, e=prob["meta"]["e"], theta0=prob["meta"]["theta0"])
        A, B = G.elliptic_stage_matrices(gen["e"], gen["theta0"], 14, gen["T"], 8)
        opts = dict(opts, max_iter=400)
    elif kind == "cw_zoh":
        prob, opts = P.cfg3_lowthrust_soc(batch=130, N=16, seed=42)
        gen = dict(kind=kind, T=2 * np.pi / 16)
        A, B = G.cw_stage_matrices(kind, 16, gen["T"])
        opts = dict(opts, max_iter=500)
    else:
        prob, opts = P.cfg2_cw_batch(batch=130, N=16, seed=43)
        gen = dict(kind=kind, T=2 * np.pi / 16)
        A, B = G.cw_stage_matrices(kind, 16, gen["T"])
        opts = dict(opts, max_iter=500)
    ref_prob = dict(prob, A=A, B=B)
    ref = cpu_oracle.solve(ref_prob, opts)
    gen_prob = {k: v for k, v in prob.items() if k not in ("A", "B", "meta")}
    gen_prob["N"] = ref_prob["A"].shape[1]
    got = solver.solve_generated(gen_prob, gen, opts)
    assert_bit_identical(got, ref, f"generated {kind}")
    assert_bit_identical(solver.solve(ref_prob, opts), ref, f"uploaded {kind}")


def test_generator_bad_arguments(solver, pkg, P):
    prob, opts = P.cfg2_cw_batch(batch=8, N=5, seed=1)
    gp = {k: v for k, v in prob.items() if k not in ("A", "B")}
    gp["N"] = 5
    for gen in (dict(kind="elliptic_zoh", T=0.1), dict(kind="cw_zoh", T=-1.0),
                dict(kind="elliptic_zoh", T=0.1, e=np.full(8, 1.5), theta0=np.zeros(8))):
        with pytest.raises(pkg._lib.AdmmError) as ei:
            solver.solve_generated(gp, gen, opts)
        assert ei.value.code == pkg._lib.E_BADARG


# ---- parallel-in-time kernel: opts.kernel = "pint" -----------------------------------------------------
# FP64, but not the oracle's operation order (chunked sweeps joined by superposition, norms summed per chunk first), so
# the bar is the north_star's: x, z, u within 1e-9 relative, same iteration count at convergence.  A problem whose
# residual sits within rounding of its threshold may stop one iteration apart; the tests allow that on at most 1 % of
# the problems (none observed on these seeds) and compare the iterates of the others.
PINT_TOL = 1e-9


def assert_pint_matches(got, ref, what):
    ig, ir = got[3]["iters"].astype(np.int64), ref[3]["iters"].astype(np.int64)
    eq = ig == ir
    assert np.abs(ig - ir).max() <= 1, f"{what}: iteration counts differ by more than one"
    assert eq.mean() >= 0.99, f"{what}: {int((~eq).sum())} of {len(eq)} iteration counts differ"
    assert np.array_equal(got[3]["status"][eq], ref[3]["status"][eq]), f"{what}: statuses differ"
    for a, b, name in zip(got[:3], ref[:3], "xzu"):
        scale = max(1.0, np.abs(b[eq]).max())
        assert np.abs(a[eq] - b[eq]).max() <= PINT_TOL * scale, f"{what}: {name} differs by {np.abs(a[eq] - b[eq]).max():.2e}"
    for k in ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho"):
        # (residuals of a problem converged down to rounding level are themselves rounding noise: absolute floor)
        assert np.allclose(got[3][k][eq], ref[3][k][eq], rtol=1e-6, atol=1e-10), f"{what}: final {k} differs"


@pytest.mark.parametrize("N", [1, 2, 3, 7, 8, 9, 20, 50])
def test_pint_kernel_matches_oracle_over_horizons(solver, cpu_oracle, P, N):
    """Every chunking: fewer stages than warps (N < 8), one stage per warp, ragged chunks (N = 9, 20, 50)."""
    prob, opts = P.cfg2_cw_batch(batch=97, N=N, seed=10 + N)
    opts = dict(opts, max_iter=3000)
    assert_pint_matches(solver.solve(prob, dict(opts, kernel="pint")), cpu_oracle.solve(prob, opts), f"pint cfg2 N={N}")


@pytest.mark.parametrize("batch", [1, 31, 33, 257])
def test_pint_kernel_ragged_batches_and_short_launches(solver, cpu_oracle, P, batch):
    prob, opts = P.cfg2_cw_batch(batch=batch, N=20, seed=3)
    opts = dict(opts, max_iter=2500)
    ref = cpu_oracle.solve(prob, opts)
    assert_pint_matches(solver.solve(prob, dict(opts, kernel="pint")), ref, f"pint batch={batch}")
    assert_pint_matches(solver.solve(prob, dict(opts, kernel="pint", chunk=7)), ref, f"pint batch={batch}, 7-iteration launches")


def test_pint_kernel_soc_adaptive_history_and_generated_model(solver, cpu_oracle, P):
    from oracle import gen_ocp as G
    prob, opts = P.cfg3_lowthrust_soc(batch=130, N=40, seed=5)            # SOC prox, zero-order-hold B
    opts = dict(opts, max_iter=3000)
    assert_pint_matches(solver.solve(prob, dict(opts, kernel="pint")), cpu_oracle.solve(prob, opts), "pint cfg3")
    prob, opts = P.cfg5_montecarlo(batch=200, N=20, seed=6)               # adaptive rho: dual rescaling inside the launch
    opts = dict(opts, max_iter=4000, adapt_every=10, adapt_until=200, adapt_mu=5.0)
    assert_pint_matches(solver.solve(prob, dict(opts, kernel="pint")), cpu_oracle.solve(prob, opts), "pint cfg5 adaptive")
    prob, opts = P.cfg1_single_impulsive()                                # residual history
    opts = dict(opts, max_iter=3000, history=1)
    got, ref = solver.solve(prob, dict(opts, kernel="pint")), cpu_oracle.solve(prob, opts)
    assert_pint_matches(got, ref, "pint cfg1")
    assert np.allclose(got[3]["hist"]["r_norm"], ref[3]["hist"]["r_norm"], rtol=1e-6, atol=1e-10, equal_nan=True)
    prob, opts = P.cfg2_cw_batch(batch=64, N=16, seed=43)                 # model generated on the device
    A, B = G.cw_stage_matrices("cw_impulsive", 16, 2 * np.pi / 16)
    gp = {k: v for k, v in prob.items() if k not in ("A", "B")}
    gp["N"] = 16
    got = solver.solve_generated(gp, dict(kind="cw_impulsive", T=2 * np.pi / 16), dict(opts, max_iter=500, kernel="pint"))
    assert_pint_matches(got, cpu_oracle.solve(dict(prob, A=A, B=B), dict(opts, max_iter=500)), "pint generated")


def test_pint_kernel_multi_tile_benchmark_shape(solver, cpu_oracle, P):
    """More tiles than SMs (two passes per CTA), N = 50, fixed 300 iterations + repacking to tolerance on a slice."""
    prob, opts = P.cfg2_cw_batch(batch=6000, N=50, seed=77)
    o = dict(opts, max_iter=300)
    assert_pint_matches(solver.solve(prob, dict(o, kernel="pint")), cpu_oracle.solve(prob, o), "pint 6000 x 50")
    sl = {k: (v[:512] if isinstance(v, np.ndarray) and v.shape[0] == 6000 else v) for k, v in prob.items()}
    assert_pint_matches(solver.solve(sl, dict(opts, kernel="pint")), cpu_oracle.solve(sl, opts), "pint 512 x 50 to tolerance")


def test_pint_kernel_falls_back_when_not_applicable(solver, cpu_oracle, P):
    """Per-problem models / quadratic cost: the pinned variant does not apply and the automatic (bit-exact) choice runs."""
    prob, opts = P.lqr_tracking(batch=40, N=10, seed=2)
    opts = dict(opts, max_iter=150)
    assert_bit_identical(solver.solve(prob, dict(opts, kernel="pint")), cpu_oracle.solve(prob, opts), "pint fallback lqr")
