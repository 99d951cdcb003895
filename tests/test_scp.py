"""sequential convex programming around the batched ADMM solve.

CPU part (-m "not gpu"): pins the SCP oracle (oracle/scp_ocp.py) -- the reference has no SCP loop, no tests and no
golden vectors (/root/reference/README.md:1-2), so the anchors are independent ones: scipy's solve_ivp on the same
nonlinear equations, finite differences for the Jacobians, the Clohessy-Wiltshire limit, the committed fixtures
(tests/golden/scp, scripts/make_golden_scp.py), and the physical check that the converged controls, flown through the
NONLINEAR dynamics, reach the target while the controls of the Clohessy-Wiltshire problem do not.

GPU part (-m gpu): `admmb_scp_solve` / `admmb_k_scp_linearise` through the C ABI against the oracle, BIT FOR BIT (the
stage records, every pass's step, the passes made, the ADMM iteration totals and the final x, z, u)."""
import glob
import os

import numpy as np
import pytest

from oracle import scp_ocp

SCP_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "scp", "*.npz")))


def _nl_rhs(t, s, a, n, R0):
    mu = n * n * R0 ** 3
    rho = np.array([R0 + s[0], s[1], s[2]])
    d3 = np.linalg.norm(rho) ** 3
    acc = np.array([2 * n * s[4] + n * n * rho[0], -2 * n * s[3] + n * n * rho[1], 0.0]) - mu * rho / d3 + a
    return np.concatenate([s[3:6], acc])


def _case(B=7, seed=5, scale=30.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    s = scale * np.array([1.0, -5.0, 0.5, 0.02, 0.01, -0.01])[None, :] * (1.0 + 0.3 * rng.standard_normal((B, 6)))
    a = 0.5 * rng.standard_normal((B, 3))
    return s, a


def test_stage_map_matches_solve_ivp():
    """The RK4 stage map converges to the exact flow of the nonlinear relative dynamics (written here with plain
    norm()/mu, not the oracle's cancellation-free form)."""
    from scipy.integrate import solve_ivp
    s, a = _case()
    T, R0, n = 0.3, 7000.0, 1.1
    F, _, _, _ = scp_ocp.linearise_stage(s, a, T, 64, n, R0)
    for i in range(s.shape[0]):
        sol = solve_ivp(_nl_rhs, (0.0, T), s[i], args=(a[i], n, R0), rtol=1e-12, atol=1e-12, method="DOP853")
        assert np.abs(F[i] - sol.y[:, -1]).max() <= 1e-8 * (1.0 + np.abs(F[i]).max())
    # fourth order: halving the step divides the error by ~16
    errs = []
    Fx, _, _, _ = scp_ocp.linearise_stage(s, a, T, 256, n, R0)
    for sub in (2, 4):
        Fs, _, _, _ = scp_ocp.linearise_stage(s, a, T, sub, n, R0)
        errs.append(np.abs(Fs - Fx).max())
    assert 10.0 < errs[0] / errs[1] < 24.0


def test_stage_jacobians_match_finite_differences_and_defect():
    s, a = _case(B=5, seed=8)
    T, R0, n, sub = 0.25, 7000.0, 1.0, 6
    F, A, Bm, c = scp_ocp.linearise_stage(s, a, T, sub, n, R0)
    h = 1e-4
    for j in range(6):
        e = np.zeros(6); e[j] = h
        Fp, *_ = scp_ocp.linearise_stage(s + e, a, T, sub, n, R0)
        Fm, *_ = scp_ocp.linearise_stage(s - e, a, T, sub, n, R0)
        assert np.abs((Fp - Fm) / (2 * h) - A[:, :, j]).max() <= 1e-7
    for j in range(3):
        e = np.zeros(3); e[j] = h
        Fp, *_ = scp_ocp.linearise_stage(s, a + e, T, sub, n, R0)
        Fm, *_ = scp_ocp.linearise_stage(s, a - e, T, sub, n, R0)
        assert np.abs((Fp - Fm) / (2 * h) - Bm[:, :, j]).max() <= 1e-7
    # the affine term closes the linearisation at the reference point
    lin = np.einsum("bij,bj->bi", A, s) + np.einsum("bij,bj->bi", Bm, a) + c
    assert np.abs(lin - F).max() <= 1e-9 * np.abs(F).max()


def test_linearisation_tends_to_clohessy_wiltshire(P):
    """Close to the chief (|r| / R0 ~ 1e-9) the stage record is the CW zero-order-hold model."""
    T = 2.0 * np.pi / 40
    Phi, Gam = P.cw_zoh(T)
    s = np.array([[1e-5, -2e-5, 1e-5, 0.0, 0.0, 0.0]])
    _, A, Bm, c = scp_ocp.linearise_stage(s, np.zeros((1, 3)), T, 32, 1.0, 7000.0)
    assert np.abs(A[0] - Phi).max() <= 1e-7
    assert np.abs(Bm[0] - Gam).max() <= 1e-7
    assert np.abs(c).max() <= 1e-12


def test_scp_controls_reach_the_target_in_the_nonlinear_dynamics(P, cpu_oracle):
    """What the outer loop is for: the converged controls, flown through the nonlinear dynamics, end at the target to
    the ADMM tolerance; the controls of the one-shot Clohessy-Wiltshire problem miss it by kilometres."""
    B, N = 6, 20
    prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=3, scale=40.0)
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    assert (info["scp_status"] == 0).all() and (info["status"] == 0).all()
    assert info["passes"].min() >= 3                        # the nonlinearity matters at this distance
    # the trajectory settles: the last step is at the level of the ADMM tolerance, far below the first
    assert (info["step"] < 1e-5 * info["hist_step"][:, 0]).all()
    ctrl = z[:, :9 * N].reshape(B, N, 9)[:, :, 6:9]
    miss = np.abs(scp_ocp.propagate_nonlinear(prob["s0"], ctrl, N, scp)).max(axis=1)
    assert miss.max() <= 5e-3
    # the same blocks on the CW model
    Phi, Gam = P.cw_zoh(scp["T"])
    cw = dict(prob, A=np.broadcast_to(Phi, (1, N, 6, 6)).copy(), B=np.broadcast_to(Gam, (1, N, 6, 3)).copy())
    _, zc, _, hc = cpu_oracle.solve(cw, opts)
    assert (hc["status"] == 0).all()
    ctrl_cw = zc[:, :9 * N].reshape(B, N, 9)[:, :, 6:9]
    miss_cw = np.abs(scp_ocp.propagate_nonlinear(prob["s0"], ctrl_cw, N, scp)).max(axis=1)
    assert miss_cw.min() >= 100.0 * miss.max()


def _fixture_workload(P, g):
    make = (P.scp_nonlinear_impulsive if str(g["control"]) == "impulsive" else
            P.scp_nonlinear_elliptic if str(g["model"]) == "nl_elliptic" else P.scp_nonlinear_rendezvous)
    return make(int(g["batch"]), int(g["N"]), seed=int(g["seed"]), scale=float(g["scale"]), substeps=int(g["substeps"]))


def test_elliptic_model_limits(P):
    """model = "nl_elliptic": (i) close to the chief the stage records are the linear elliptic STMs of the generator
    oracle (config 4's model, itself checked against solve_ivp in tests/test_oracle.py); (ii) with e = 0 it is the
    circular model; (iii) the true-anomaly table advances by the orbit's own clock (Kepler's equation)."""
    from oracle import gen_ocp
    rng = np.random.Generator(np.random.PCG64(1))
    B, N = 4, 6
    e, th0, T = rng.uniform(0.05, 0.5, B), rng.uniform(0.0, 6.0, B), 2.0 * np.pi / N
    scp = dict(model="nl_elliptic", T=T, R0=7000.0, substeps=8, e=e, theta0=th0)
    Ag, Bg = gen_ocp.elliptic_stage_matrices(e, th0, N, T, 8)
    A, Bm, c = scp_ocp.linearise(np.full((B, 9 * N + 6), 1e-7), N, scp)
    assert np.abs(A - Ag).max() <= 1e-7 and np.abs(Bm - Bg).max() <= 1e-7 and np.abs(c).max() <= 1e-12
    xr = 30.0 * rng.standard_normal((B, 9 * N + 6))
    A0, B0, c0 = scp_ocp.linearise(xr, N, dict(scp, e=np.zeros(B)))
    Ac, Bc, cc = scp_ocp.linearise(xr, N, dict(model="nl_circular", T=T, R0=7000.0, substeps=8))
    assert np.abs(A0 - Ac).max() <= 1e-13 and np.abs(B0 - Bc).max() <= 1e-13 and np.abs(c0 - cc).max() <= 1e-11
    # one full period brings the true anomaly back (mod 2 pi): N stages of 2 pi / N at mean motion 1
    tab = scp_ocp.theta_table(e, th0, N + 1, T, 64)
    assert np.abs((tab[:, N] - th0) - 2.0 * np.pi).max() <= 1e-7


def test_scp_elliptic_controls_reach_the_target(P):
    B, N = 4, 14
    prob, scp, opts = P.scp_nonlinear_elliptic(B, N, seed=3, scale=30.0)
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    assert (info["scp_status"] == 0).all() and (info["status"] == 0).all() and info["passes"].min() >= 3
    ctrl = z[:, :9 * N].reshape(B, N, 9)[:, :, 6:9]
    miss = np.abs(scp_ocp.propagate_nonlinear(prob["s0"], ctrl, N, scp)).max(axis=1)
    assert miss.max() <= 5e-3


def test_scp_golden_fixtures_exist():
    assert len(SCP_GOLDEN) >= 4


def test_impulsive_stage_is_a_coast_after_the_velocity_increment():
    """control = "impulsive": s+ = F(s + [0; dv]); A = dF/ds at the post-impulse state, B = A[:, 3:6]; checked against
    solve_ivp of the coast and finite differences in dv."""
    from scipy.integrate import solve_ivp
    s, dv = _case(B=4, seed=11)
    T, R0, n, sub = 0.3, 7000.0, 1.0, 48
    F, A, Bm, c = scp_ocp.linearise_stage(s, dv, T, sub, n, R0, impulsive=True)
    assert np.array_equal(Bm, A[:, :, 3:6])
    for i in range(s.shape[0]):
        s_in = s[i].copy(); s_in[3:] += dv[i]
        sol = solve_ivp(_nl_rhs, (0.0, T), s_in, args=(np.zeros(3), n, R0), rtol=1e-12, atol=1e-12, method="DOP853")
        assert np.abs(F[i] - sol.y[:, -1]).max() <= 1e-8 * (1.0 + np.abs(F[i]).max())
    h = 1e-4
    for j in range(3):
        e = np.zeros(3); e[j] = h
        Fp, *_ = scp_ocp.linearise_stage(s, dv + e, T, sub, n, R0, impulsive=True)
        Fm, *_ = scp_ocp.linearise_stage(s, dv - e, T, sub, n, R0, impulsive=True)
        assert np.abs((Fp - Fm) / (2 * h) - Bm[:, :, j]).max() <= 1e-7
    lin = np.einsum("bij,bj->bi", A, s) + np.einsum("bij,bj->bi", Bm, dv) + c
    assert np.abs(lin - F).max() <= 1e-9 * np.abs(F).max()


def test_scp_impulsive_controls_reach_the_target(P):
    B, N = 4, 12
    prob, scp, opts = P.scp_nonlinear_impulsive(B, N, seed=9, scale=30.0)
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    assert (info["scp_status"] == 0).all() and (info["status"] == 0).all()
    dv = z[:, :9 * N].reshape(B, N, 9)[:, :, 6:9]
    s = prob["s0"].copy()
    for k in range(N):
        s, *_ = scp_ocp.linearise_stage(s, dv[:, k], scp["T"], scp["substeps"], 1.0, scp["R0"], impulsive=True)
    assert np.abs(s).max() <= 5e-3
    assert np.abs(dv).max() <= prob["block_par"][0, 2, 5] + 1e-9          # box on every increment


@pytest.mark.parametrize("path", SCP_GOLDEN, ids=[os.path.basename(p) for p in SCP_GOLDEN])
def test_scp_oracle_reproduces_golden_fixtures(P, path):
    g = np.load(path)
    prob, scp, opts = _fixture_workload(P, g)
    assert np.array_equal(prob["s0"], g["s0"])
    xref0, A0, B0, c0 = scp_ocp.shoot(prob["s0"], None, int(g["N"]), scp)
    for a, k in ((xref0, "xref0"), (A0, "A0"), (B0, "B0"), (c0, "c0")):
        assert np.array_equal(a, g[k]), k
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    for a, k in ((x, "x"), (z, "z"), (u, "u")):
        assert np.array_equal(a, g[k]), k
    for k in ("passes", "scp_status", "iters_total", "iters", "status"):
        assert np.array_equal(info[k], g[k]), k
    assert np.array_equal(info["hist_step"], g["hist_step"], equal_nan=True)


def test_scp_results_do_not_depend_on_the_batch(P):
    """Per-problem exit: a problem solved alone makes the same passes and ends on the same bits as inside a batch
    (what makes the sharded solve independent of the number of GPUs)."""
    prob, scp, opts = P.scp_nonlinear_rendezvous(5, 10, seed=12, scale=25.0, substeps=3)
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    for i in (0, 3):
        one = dict(prob, s0=prob["s0"][i:i + 1])
        x1, z1, u1, i1 = scp_ocp.scp_solve(one, scp, opts)
        assert np.array_equal(x1[0], x[i]) and np.array_equal(u1[0], u[i])
        assert i1["passes"][0] == info["passes"][i] and i1["iters_total"][0] == info["iters_total"][i]


# ------------------------------------------------------------------------------------------------------------------
# GPU: the device loop against the oracle, bit for bit
# ------------------------------------------------------------------------------------------------------------------
def _assert_scp_equal(got, ref):
    xg, zg, ug, ig = got
    xr, zr, ur, ir = ref
    assert np.array_equal(ig["passes"], ir["passes"]), f"passes differ: {ig['passes']} vs {ir['passes']}"
    assert np.array_equal(ig["hist_step"], ir["hist_step"], equal_nan=True), "per-pass steps differ"
    for k in ("scp_status", "iters_total", "iters", "status"):
        assert np.array_equal(ig[k], ir[k]), k
    assert np.array_equal(ig["step"], ir["step"], equal_nan=True)
    for a, b, name in ((xg, xr, "x"), (zg, zr, "z"), (ug, ur, "u")):
        assert np.array_equal(a, b), f"{name} differs, max |d| = {np.nanmax(np.abs(a - b)):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,sub", [(1, 3, 2), (33, 7, 3), (130, 12, 4)])
def test_scp_linearise_kernels_bit_identical(solver, B, N, sub):
    rng = np.random.Generator(np.random.PCG64(100 + B))
    scp = dict(T=2.0 * np.pi / N, R0=6900.0, nmm=1.05, substeps=sub)
    n = 9 * N + 6
    xref = 40.0 * rng.standard_normal((B, n))
    A, Bm, c, _ = solver.k_scp_linearise(N, scp, xref)
    Ao, Bo, co = scp_ocp.linearise(xref, N, scp)
    assert np.array_equal(A, Ao) and np.array_equal(Bm, Bo) and np.array_equal(c, co)
    # along the nonlinear trajectory from s0 under given controls
    s0 = 30.0 * rng.standard_normal((B, 6))
    ctrl = rng.standard_normal((B, N, 3))
    xin = np.zeros((B, n))
    xin[:, :9 * N].reshape(B, N, 9)[:, :, 6:9] = ctrl
    A, Bm, c, xr = solver.k_scp_linearise(N, scp, xin, s0=s0)
    xo, Ao, Bo, co = scp_ocp.shoot(s0, ctrl, N, scp)
    assert np.array_equal(xr, xo)
    assert np.array_equal(A, Ao) and np.array_equal(Bm, Bo) and np.array_equal(c, co)


@pytest.mark.gpu
def test_scp_impulsive_linearise_and_solve_bit_identical(solver, P):
    rng = np.random.Generator(np.random.PCG64(77))
    B, N = 37, 9
    scp = dict(T=2.0 * np.pi / N, R0=6900.0, nmm=1.0, substeps=3, control="impulsive")
    xref = 40.0 * rng.standard_normal((B, 9 * N + 6))
    A, Bm, c, _ = solver.k_scp_linearise(N, scp, xref)
    Ao, Bo, co = scp_ocp.linearise(xref, N, scp)
    assert np.array_equal(A, Ao) and np.array_equal(Bm, Bo) and np.array_equal(c, co)
    assert np.array_equal(Bm, A[:, :, :, 3:6])
    prob, scp, opts = P.scp_nonlinear_impulsive(B, 12, seed=9, scale=30.0, substeps=3)
    _assert_scp_equal(solver.scp_solve(prob, scp, opts), scp_ocp.scp_solve(prob, scp, opts))


@pytest.mark.gpu
@pytest.mark.parametrize("control", ["zoh", "impulsive"])
def test_scp_elliptic_linearise_and_solve_bit_identical(solver, P, control):
    rng = np.random.Generator(np.random.PCG64(78))
    B, N = 41, 7
    scp = dict(model="nl_elliptic", control=control, T=2.0 * np.pi / N, R0=6900.0, substeps=3,
               e=rng.uniform(0.0, 0.7, B), theta0=rng.uniform(-5.0, 5.0, B))
    xref = 40.0 * rng.standard_normal((B, 9 * N + 6))
    A, Bm, c, _ = solver.k_scp_linearise(N, scp, xref)
    Ao, Bo, co = scp_ocp.linearise(xref, N, scp)
    assert np.array_equal(A, Ao) and np.array_equal(Bm, Bo) and np.array_equal(c, co)
    s0 = 30.0 * rng.standard_normal((B, 6))
    A, Bm, c, xr = solver.k_scp_linearise(N, scp, np.zeros((B, 9 * N + 6)), s0=s0)
    xo, Ao, Bo, co = scp_ocp.shoot(s0, None, N, scp)
    assert np.array_equal(xr, xo) and np.array_equal(A, Ao) and np.array_equal(Bm, Bo) and np.array_equal(c, co)
    if control == "zoh":
        prob, scp, opts = P.scp_nonlinear_elliptic(B, 12, seed=33, scale=25.0, substeps=3)
        _assert_scp_equal(solver.scp_solve(prob, scp, opts), scp_ocp.scp_solve(prob, scp, opts))


@pytest.mark.gpu
@pytest.mark.parametrize("path", SCP_GOLDEN, ids=[os.path.basename(p) for p in SCP_GOLDEN])
def test_scp_gpu_reproduces_golden_fixtures(solver, P, path):
    g = np.load(path)
    prob, scp, opts = _fixture_workload(P, g)
    x, z, u, info = solver.scp_solve(prob, scp, opts)
    ref = (g["x"], g["z"], g["u"], {k: g[k] for k in ("passes", "hist_step", "scp_status", "iters_total", "iters", "status", "step")})
    _assert_scp_equal((x, z, u, info), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(1, 8), (37, 10), (96, 20)])
def test_scp_solve_matches_oracle(solver, P, B, N):
    prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=20 + B, scale=30.0, substeps=3)
    got = solver.scp_solve(prob, scp, opts)
    ref = scp_ocp.scp_solve(prob, scp, opts)
    _assert_scp_equal(got, ref)
    info = got[3]
    assert info["scp_stats"][0] == int((ref[3]["scp_status"] == 0).sum())
    assert info["scp_stats"][1] == int(ref[3]["iters_total"].sum()) == info["stats"][1]
    assert info["scp_stats"][2] == int(ref[3]["passes"].max()) and info["scp_stats"][3] == int(ref[3]["passes"].sum())
    assert info["passes"].max() > info["passes"].min() or B == 1      # the per-problem exit is exercised
    assert info["linearise_ms"] > 0.0 and info["launches"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(2, 1), (3, 2), (31, 3), (33, 5)])
def test_scp_tiny_horizons_and_ragged_batches(solver, P, B, N):
    """Horizons shorter than the TMA ring's prefetch distance, batches that do not fill a warp."""
    prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=300 + B, scale=10.0, substeps=2, max_pass=6)
    _assert_scp_equal(solver.scp_solve(prob, scp, opts), scp_ocp.scp_solve(prob, scp, opts))


@pytest.mark.gpu
def test_scp_per_problem_block_parameters_and_initial_warm_start(solver, P):
    """Per-problem thrust bounds / terminal points (par_batched) travel through the passes; a warm start given with the
    problem serves the first pass."""
    B, N = 20, 8
    prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=51, scale=25.0, substeps=3, max_pass=8)
    rng = np.random.Generator(np.random.PCG64(6))
    bp = np.broadcast_to(prob["block_par"], (B,) + prob["block_par"].shape[1:]).copy()
    bp[:, 2:3 * N:3, 1] *= rng.uniform(0.9, 1.3, (B, 1))                      # thrust bound per problem
    bp[:, 3 * N, 2:5] = 0.5 * rng.standard_normal((B, 3))                     # terminal position per problem
    n = 9 * N + 6
    prob = dict(prob, block_par=bp, z0=0.1 * rng.standard_normal((B, n)), u0=0.01 * rng.standard_normal((B, n)))
    got = solver.scp_solve(prob, scp, opts)
    ref = scp_ocp.scp_solve(prob, scp, opts)
    _assert_scp_equal(got, ref)
    assert (got[3]["scp_status"] == 0).all()


@pytest.mark.gpu
def test_scp_max_pass_and_quadratic_cost(solver, P):
    """max_pass cuts the loop (status 1, same bits as the oracle cut at the same pass); per-problem Q, R and a linear
    cost q travel through the passes."""
    B, N = 24, 9
    prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=41, scale=50.0, substeps=2, max_pass=3)
    rng = np.random.Generator(np.random.PCG64(4))
    prob = dict(prob, Q=np.broadcast_to(np.diag([1e-4] * 3 + [1e-3] * 3), (B, N + 1, 6, 6)).copy(),
                R=np.broadcast_to(1e-2 * np.eye(3), (B, N, 3, 3)).copy(), q=1e-3 * rng.standard_normal((B, 9 * N + 6)))
    got = solver.scp_solve(prob, scp, opts)
    ref = scp_ocp.scp_solve(prob, scp, opts)
    _assert_scp_equal(got, ref)
    assert (got[3]["scp_status"] == 1).any() and got[3]["passes"].max() == 3


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["circular", "elliptic"])
def test_scp_two_gpus_same_bits(pkg, P, solver, workload):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    make = P.scp_nonlinear_elliptic if workload == "elliptic" else P.scp_nonlinear_rendezvous     # per-problem e, theta0 are sharded too
    prob, scp, opts = make(70, 10, seed=77, scale=30.0, substeps=3)
    one = solver.scp_solve(prob, scp, opts)
    with pkg.Solver(devices=[0, 1]) as s2:
        two = s2.scp_solve(prob, scp, opts)
    _assert_scp_equal(two, one)
    assert two[3]["scp_stats"] == one[3]["scp_stats"]


@pytest.mark.gpu
def test_scp_refuses_what_it_cannot_carry_across_passes(solver, pkg, P):
    prob, scp, opts = P.scp_nonlinear_rendezvous(4, 6, seed=1)
    for bad_opts in (dict(opts, adapt_rho=1), dict(opts, history=1), dict(opts, xupdate="dense")):
        with pytest.raises(pkg._lib.AdmmError) as ei:
            solver.scp_solve(prob, scp, bad_opts)
        assert ei.value.code == pkg._lib.E_BADARG
    for bad_scp in (dict(scp, R0=0.0), dict(scp, max_pass=0), dict(scp, T=-1.0)):
        with pytest.raises(pkg._lib.AdmmError) as ei:
            solver.scp_solve(prob, bad_scp, opts)
        assert ei.value.code == pkg._lib.E_BADARG
    # the handle is still good
    x, z, u, info = solver.scp_solve(prob, scp, opts)
    assert (info["scp_status"] == 0).all()
