"""CPU tests (-m "not gpu"): pin the oracle.  The reference ships no tests or golden vectors
(README + LICENSE only), so the oracle is anchored by independent checks:
KKT residuals, dense-KKT vs Riccati agreement, prox identities, an LP solved by HiGHS, expm checks of
the dynamics, and agreement between the NumPy and the C restatements (incl. the committed fixtures)."""
import glob
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy.linalg import expm
from scipy.optimize import linprog

from oracle import admm_ocp as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


# ----------------------------------------------------------------------------- dynamics (T1)
def test_cw_stm_matches_expm(P):
    for T in (0.05, 2 * np.pi / 50, 1.3):
        assert np.allclose(P.cw_stm(T), expm(P.cw_continuous() * T), atol=1e-12)


def test_cw_zoh_matches_augmented_expm(P):
    T = 2 * np.pi / 20
    M = np.zeros((9, 9))
    M[:6, :6] = P.cw_continuous()
    M[3:6, 6:9] = np.eye(3)
    E = expm(M * T)
    Phi, Gam = P.cw_zoh(T)
    assert np.allclose(Phi, E[:6, :6], atol=1e-12)
    assert np.allclose(Gam, E[:6, 6:9], atol=1e-12)


def test_elliptic_e0_reduces_to_cw(P):
    A, B = P.elliptic_stage_matrices(np.array([0.0, 0.0]), np.array([0.3, 2.0]), N=4, T=0.2, substeps=16)
    Phi, Gam = P.cw_zoh(0.2)
    assert np.allclose(A, Phi[None, None], atol=1e-8)
    assert np.allclose(B, Gam[None, None], atol=1e-8)


# ----------------------------------------------------------------------------- x-update (T0)
def _random_model(rng, N, with_cost=True, with_c=True):
    A = np.eye(6)[None, None] + 0.3 * rng.standard_normal((1, N, 6, 6))
    B = rng.standard_normal((1, N, 6, 3))
    c = 0.1 * rng.standard_normal((1, N, 6)) if with_c else None
    Q = R = None
    if with_cost:
        q0 = rng.standard_normal((1, N + 1, 6, 6))
        Q = q0 @ np.swapaxes(q0, -1, -2)
        r0 = rng.standard_normal((1, N, 3, 3))
        R = r0 @ np.swapaxes(r0, -1, -2)
    return A, B, c, Q, R


@pytest.mark.parametrize("with_cost,with_c,states_split", [(True, True, True), (False, False, False),
                                                          (True, False, False), (False, True, True)])
def test_riccati_xupdate_solves_the_kkt_system(with_cost, with_c, states_split):
    rng = np.random.default_rng(1)
    N, rho = 7, 0.7
    n = 9 * N + 6
    A, B, c, Q, R = _random_model(rng, N, with_cost, with_c)
    bt = np.full(3 * N + 2, O.BLK_FREE if states_split else O.BLK_NONE)
    bt[2::3] = O.BLK_BOX
    bt[3 * N:] = O.BLK_POINT
    wblk, w = O.split_weights(bt)
    fac = O.riccati_factor(A, B, c, Q, R, rho, wblk)
    s0 = rng.standard_normal((3, 6))
    rt = w * rng.standard_normal((3, n))
    x = O.xupdate_riccati(fac, A, B, c, s0, rt)
    dfac = O.kkt_dense_factor(A[0], B[0], None if c is None else c[0], None if Q is None else Q[0],
                              None if R is None else R[0], rho, wblk)
    G = O.assemble_G(A[0], B[0], N)
    for p in range(3):
        h = np.concatenate([s0[p], np.zeros(6 * N) if c is None else c[0].reshape(-1)])
        sol = np.linalg.solve(dfac["KKT"], np.concatenate([rt[p], h]))
        assert np.allclose(x[p], sol[:n], atol=1e-9)
        assert np.abs(G @ x[p] - h).max() < 1e-10              # primal feasibility of the dynamics
    xd = O.xupdate_dense(dfac, s0, rt)                          # a2' == a2
    assert np.allclose(xd, x, atol=1e-9)


# ----------------------------------------------------------------------------- prox identities (T0)
vec3 = st.lists(st.floats(-5, 5, allow_nan=False, width=64), min_size=3, max_size=3)


def _prox(t, v, lam=0.7, rad=1.3, lo=(-0.5, -1.0, 0.2), hi=(0.5, 0.0, 0.9), rinv=0.6):
    par = np.zeros((1, 1, 8))
    par[0, 0] = [lam, rad, *lo, *hi]
    return O.prox_blocks(np.asarray(v, dtype=float)[None, :], np.array([t]), par, np.array([rinv]))[0]


@settings(max_examples=150, deadline=None)
@given(vec3, vec3)
def test_prox_nonexpansive_and_projections_idempotent(v, w):
    v, w = np.array(v), np.array(w)
    for t in range(8):
        pv, pw = _prox(t, v), _prox(t, w)
        assert np.linalg.norm(pv - pw) <= np.linalg.norm(v - w) + 1e-12
    for t in (O.BLK_BOX, O.BLK_BALL, O.BLK_POINT, O.BLK_FREE):
        pv = _prox(t, v)
        assert np.allclose(_prox(t, pv), pv, atol=1e-12)


@settings(max_examples=150, deadline=None)
@given(vec3)
def test_moreau_decomposition_of_the_norm_proxes(v):
    v = np.array(v)
    kap = 0.7 * 0.6
    assert np.allclose(v - _prox(O.BLK_L1, v), np.clip(v, -kap, kap), atol=1e-12)
    nv = np.linalg.norm(v)
    dual = v if nv <= kap else v * kap / nv
    assert np.allclose(v - _prox(O.BLK_L2, v), dual, atol=1e-12)


def test_l1_box_prox_is_the_1d_argmin():
    grid = np.linspace(-1.5, 1.5, 60001)
    for v in (-2.0, -0.3, 0.1, 0.45, 0.8, 3.0):
        z = _prox(O.BLK_L1_BOX, [v, v, v], lo=(-0.5, -0.5, -0.5), hi=(0.5, 0.5, 0.5))[0]
        g = grid[(grid >= -0.5) & (grid <= 0.5)]
        best = g[np.argmin(0.7 * 0.6 * np.abs(g) + 0.5 * (g - v) ** 2)]
        assert abs(z - best) < 1e-4


def test_l2_ball_prox_is_radial_shrink_then_clip():
    v = np.array([3.0, -4.0, 12.0])                      # |v| = 13
    z = _prox(O.BLK_L2_BALL, v, lam=1.0, rad=2.0, rinv=1.0)
    assert np.allclose(z, v * 2.0 / 13.0)
    z = _prox(O.BLK_L2_BALL, v, lam=1.0, rad=20.0, rinv=1.0)
    assert np.allclose(z, v * 12.0 / 13.0)
    assert np.allclose(_prox(O.BLK_L2_BALL, [0.1, 0.1, 0.1], lam=1.0, rad=2.0, rinv=1.0), 0.0)


# ----------------------------------------------------------------------------- independent solver (T1)
def test_cfg1_objective_matches_highs_lp(P, cpu_oracle):
    prob, opts = P.cfg1_single_impulsive()
    N = prob["N"]
    n = 9 * N + 6
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=20000, abstol=1e-8, reltol=1e-8))
    assert h["status"][0] == 0
    G = O.assemble_G(prob["A"][0], prob["B"][0], N)
    hvec = np.zeros(6 * (N + 1))
    hvec[:6] = prob["s0"][0]
    idx = [9 * k + 6 + i for k in range(N) for i in range(3)]
    nt = len(idx)
    cost = np.concatenate([np.zeros(n), np.ones(nt)])
    Et = np.zeros((6, n + nt))
    Et[:, 9 * N:9 * N + 6] = np.eye(6)
    Aeq = np.vstack([np.hstack([G, np.zeros((G.shape[0], nt))]), Et])
    beq = np.concatenate([hvec, np.zeros(6)])
    Aub = np.zeros((2 * nt, n + nt))
    for j, i in enumerate(idx):
        Aub[2 * j, i], Aub[2 * j, n + j] = 1, -1
        Aub[2 * j + 1, i], Aub[2 * j + 1, n + j] = -1, -1
    bounds = [(None, None)] * n + [(0, None)] * nt
    for i in idx:
        bounds[i] = (-0.4, 0.4)
    lp = linprog(cost, A_ub=Aub, b_ub=np.zeros(2 * nt), A_eq=Aeq, b_eq=beq, bounds=bounds, method="highs")
    assert lp.status == 0
    assert abs(O.objective(prob, z)[0] - lp.fun) < 1e-5 * max(1.0, abs(lp.fun))
    assert np.abs(G @ x[0] - hvec).max() < 1e-9


def test_converged_point_satisfies_optimality_conditions_soc(P, cpu_oracle):
    """No SOCP solver in the image: check stationarity directly.  At a fixed point, rho*u is a
    subgradient of g at z and P x + q + G' nu + rho*u = 0 for some nu (i.e. the residual projected on
    the null space of G vanishes)."""
    prob, opts = P.cfg3_lowthrust_soc(batch=2, N=12, seed=5)
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=100000, abstol=1e-9, reltol=1e-9))
    N = 12
    G = O.assemble_G(prob["A"][0], prob["B"][0], N)
    _, w = O.split_weights(prob["block_type"])
    for p in range(2):
        if h["status"][p] != 0:
            continue
        grad = h["rho"][p] * w * u[p]
        nu = np.linalg.lstsq(G.T, -grad, rcond=None)[0]
        assert np.abs(G.T @ nu + grad).max() < 1e-5
        zb = z[p].reshape(-1, 3)
        for k in range(N):
            assert np.linalg.norm(zb[3 * k + 2]) <= prob["block_par"][0, 3 * k + 2, 1] + 1e-9


@pytest.mark.parametrize("elliptic", [False, True])
def test_soc_objective_matches_an_independent_nlp_solver(P, cpu_oracle, elliptic):
    """the thrust-magnitude (second-order-cone) configurations had no independent solver behind
    them.  The image has no SOCP solver, but the condensed problem is small and convex:
        min sum_k lam |a_k|_2   s.t.  sum_k G_k a_k = -Phi s0 (terminal point),  |a_k|_2 <= a_max,
    solved here by scipy's SLSQP with the norm smoothed to sqrt(|a|^2 + eps^2) (eps = 1e-7: the epigraph form t^2 >= |a|^2
    loses its constraint qualification where a control is exactly zero, and fuel-optimal controls are), for the shared
    CW model (config 3) and a per-problem elliptic model (config 4).  The ADMM optimum must (i) be feasible, (ii) not
    exceed the value of SLSQP's feasible point, (iii) agree with it to SLSQP's accuracy."""
    from scipy.optimize import minimize
    N, B, eps = 10, 3, 1e-7
    prob, opts = (P.cfg4_elliptic(batch=B, N=N, seed=8) if elliptic else P.cfg3_lowthrust_soc(batch=B, N=N, seed=5))
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=300000, abstol=1e-9, reltol=1e-9))
    assert (h["status"] == 0).all()
    lam, rad = prob["block_par"][0, 2, 0], prob["block_par"][0, 2, 1]
    for i in range(B):
        A, Bm = (prob["A"][i], prob["B"][i]) if elliptic else (prob["A"][0], prob["B"][0])
        G, M = [], np.eye(6)
        for k in range(N - 1, -1, -1):
            G.insert(0, M @ Bm[k])
            M = M @ A[k]
        Gm, rhs = np.hstack(G), -(M @ prob["s0"][i])

        def f(v):
            return lam * np.sqrt((v.reshape(N, 3) ** 2).sum(1) + eps * eps).sum()

        def grad(v):
            a = v.reshape(N, 3)
            return (lam * a / np.sqrt((a * a).sum(1) + eps * eps)[:, None]).ravel()

        def ball(v):
            return rad * rad - (v.reshape(N, 3) ** 2).sum(1)

        def ball_jac(v):
            a = v.reshape(N, 3)
            J = np.zeros((N, 3 * N))
            for k in range(N):
                J[k, 3 * k:3 * k + 3] = -2.0 * a[k]
            return J

        a0 = np.linalg.lstsq(Gm, rhs, rcond=None)[0]
        r = minimize(f, a0, jac=grad, method="SLSQP",
                     constraints=[dict(type="eq", fun=lambda v: Gm @ v - rhs, jac=lambda v: Gm),
                                  dict(type="ineq", fun=ball, jac=ball_jac)], options=dict(maxiter=5000, ftol=1e-14))
        # SLSQP may stop with 'positive directional derivative' at this tolerance: what matters is that its point is feasible
        assert np.abs(Gm @ r.x - rhs).max() <= 1e-7 and ball(r.x).min() >= -1e-8
        fun = lam * np.linalg.norm(r.x.reshape(N, 3), axis=1).sum()
        ctrl = z[i, :9 * N].reshape(N, 9)[:, 6:9]
        nrm = np.linalg.norm(ctrl, axis=1)
        obj = lam * nrm.sum()
        assert np.abs(Gm @ ctrl.ravel() - rhs).max() <= 1e-6 and nrm.max() <= rad + 1e-9      # (i)
        assert obj <= fun * (1.0 + 1e-6)                                                        # (ii)
        assert abs(obj - fun) <= 1e-5 * fun, (obj, fun)                                         # (iii)


# ----------------------------------------------------------------------------- NumPy vs C restatement
def _close(a, b, tol=1e-10):
    return np.allclose(a, b, rtol=tol, atol=tol, equal_nan=True)


@pytest.mark.parametrize("case", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "lqr", "lqr_pp_adapt", "literal", "dense"])
def test_c_oracle_matches_numpy_oracle(P, cpu_oracle, case):
    if case == "cfg1":
        prob, opts = P.cfg1_single_impulsive(); opts = dict(opts, max_iter=3000)
    elif case == "cfg2":
        prob, opts = P.cfg2_cw_batch(6, 12, seed=3); opts = dict(opts, max_iter=400)
    elif case == "cfg3":
        prob, opts = P.cfg3_lowthrust_soc(5, 14, seed=3); opts = dict(opts, max_iter=400)
    elif case == "cfg4":
        prob, opts = P.cfg4_elliptic(5, 10, seed=3); opts = dict(opts, max_iter=400)
    elif case == "cfg5":
        prob, opts = P.cfg5_montecarlo(6, 10, seed=3); opts = dict(opts, max_iter=500, adapt_every=10)
    elif case == "lqr":
        prob, opts = P.lqr_tracking(5, 9, seed=3); opts = dict(opts, max_iter=200)
    elif case == "lqr_pp_adapt":
        prob, opts = P.lqr_tracking(4, 9, seed=3, per_problem=True)
        opts = dict(opts, max_iter=200, adapt_rho=1, adapt_every=5, adapt_mu=2.0)
    elif case == "literal":
        prob, opts = P.cfg2_cw_batch(4, 10, seed=4)
        bt = prob["block_type"].copy(); bt[bt == P.BLK_NONE] = P.BLK_FREE
        prob = dict(prob, block_type=bt); opts = dict(opts, max_iter=300, alpha=1.5)
    else:
        prob, opts = P.lqr_tracking(4, 8, seed=5); opts = dict(opts, max_iter=150, xupdate="dense")
    opts = dict(opts, history=1)
    xn, zn, un, hn = O.admm_solve(prob, opts)
    xc, zc, uc, hc = cpu_oracle.solve(prob, opts)
    assert np.array_equal(hn["iters"], hc["iters"])
    assert np.array_equal(hn["status"], hc["status"])
    # the two restatements differ only in summation order / fma use (~1e-16 per operation)
    assert _close(xn, xc, 1e-8) and _close(zn, zc, 1e-8) and _close(un, uc, 1e-8)
    assert _close(hn["rho"], hc["rho"])
    assert hn["refactor_count"] == hc["refactor_count"]
    assert _close(hn["hist"]["r_norm"], hc["hist"]["r_norm"], 1e-8)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_c_oracle_reproduces_golden_fixtures(cpu_oracle, path):
    g = np.load(path)
    prob = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    prob["N"] = prob["A"].shape[1]
    opts = {k[4:]: g[k].item() for k in g.files if k.startswith("opt_")}
    x, z, u, h = cpu_oracle.solve(prob, opts)
    assert np.array_equal(h["iters"], g["out_iters"])
    assert np.array_equal(h["status"], g["out_status"])
    assert _close(x, g["out_x"], 1e-8) and _close(z, g["out_z"], 1e-8) and _close(u, g["out_u"], 1e-8)
    assert _close(h["hist"]["r_norm"], g["out_hist_r"], 1e-7)
    assert h["refactor_count"] == int(g["out_refactor"])


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 7


# ----------------------------------------------------------------------------- driver edge cases
def test_edge_cases_max_iter_one_and_status_codes(P, cpu_oracle):
    prob, opts = P.cfg2_cw_batch(3, 5, seed=1)
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=1))
    assert list(h["iters"]) == [1, 1, 1] and list(h["status"]) == [1, 1, 1]
    xn, zn, un, hn = O.admm_solve(prob, dict(opts, max_iter=1))
    assert _close(x, xn) and _close(z, zn)
    # a NaN in s0 is reported per problem, not as a failure of the call
    prob["s0"][1, 0] = np.nan
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=50))
    assert h["status"][1] == 2 and h["status"][0] != 2


def test_split_free_formulation_converges_to_same_solution_slower(P, cpu_oracle):
    prob, opts = P.cfg1_single_impulsive(N=20)
    o = dict(opts, alpha=1.0, max_iter=200000, abstol=1e-7, reltol=1e-7)
    x1, z1, u1, h1 = cpu_oracle.solve(prob, o)
    bt = prob["block_type"].copy(); bt[bt == P.BLK_NONE] = P.BLK_FREE
    x2, z2, u2, h2 = cpu_oracle.solve(dict(prob, block_type=bt), o)
    assert h1["status"][0] == 0 and h2["status"][0] == 0
    assert h2["iters"][0] > 2 * h1["iters"][0]
    assert abs(O.objective(prob, z1)[0] - O.objective(prob, z2)[0]) < 1e-4


# ---- generator oracle: oracle/gen_ocp.py is the spelled-out-order statement the device generators follow ----
GEN_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "generators", "*.npz")))


def test_det_sincos_is_accurate_and_reproduces_its_fixture():
    from oracle import gen_ocp as G
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "generators", "det_sincos.npz"))
    s, c = G.det_sincos(g["x"])
    assert np.array_equal(s, g["s"]) and np.array_equal(c, g["c"])           # same bits on every machine
    assert np.abs(s - np.sin(g["x"])).max() <= 4e-16 and np.abs(c - np.cos(g["x"])).max() <= 4e-16


def test_generator_oracle_matches_problem_generators(P):
    """gen_ocp (fixed operation order, polynomial trig) against problems.py (NumPy matmul, libm trig)."""
    from oracle import gen_ocp as G
    for T, nmm in ((2 * np.pi / 50, 1.0), (0.37, 1.3), (2 * np.pi / 20, 0.8)):
        assert np.abs(G.cw_stm(T, nmm) - P.cw_stm(T, nmm)).max() <= 1e-14
        Pg, Gg = G.cw_zoh(T, nmm)
        Pp, Gp = P.cw_zoh(T, nmm)
        assert np.abs(Pg - Pp).max() <= 1e-14 and np.abs(Gg - Gp).max() <= 1e-14
    rng = np.random.default_rng(5)
    e, th = rng.uniform(0.0, 0.7, 40), rng.uniform(0, 2 * np.pi, 40)
    A, B = G.elliptic_stage_matrices(e, th, 12, 2 * np.pi / 12, 8)
    A2, B2 = P.elliptic_stage_matrices(e, th, 12, 2 * np.pi / 12, 8)
    assert np.abs(A - A2).max() <= 1e-13 * np.abs(A2).max() and np.abs(B - B2).max() <= 1e-13 * np.abs(B2).max()


@pytest.mark.parametrize("path", GEN_GOLDEN, ids=[os.path.basename(g)[:-4] for g in GEN_GOLDEN])
def test_generator_oracle_reproduces_golden_fixtures(path):
    from oracle import gen_ocp as G
    g = np.load(path)
    if "kind" not in g.files:
        return
    kind = str(g["kind"])
    if kind == "elliptic_zoh":
        A, B = G.elliptic_stage_matrices(g["e"], g["theta0"], int(g["N"]), float(g["T"]), int(g["substeps"]))
    else:
        A, B = G.cw_stage_matrices(kind, int(g["N"]), float(g["T"]), float(g["nmm"]))
    assert np.array_equal(A, g["A"]) and np.array_equal(B, g["B"])
