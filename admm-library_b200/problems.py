"""Problem generators for the batched optimal-control QPs (host side, NumPy only).

The reference (/root/reference/README.md:1-2) states only "ADMM for astrodynamics
problems"; the problem family is fixed by BASELINE.json `north_star`/`configs` and
SURVEY.md section 8(d).  The solver takes stage matrices as *inputs*, so nothing
here is on the device hot path: these functions only build (A_k, B_k, c_k, s0,
block descriptors) for the five benchmark configurations.

Conventions (shared by the oracle, the C ABI and the MATLAB surface)
-------------------------------------------------------------------
state   s = [x, y, z, vx, vy, vz]  in LVLH (x radial, y along-track, z cross-track)
control a in R^3 (impulsive dv or thrust acceleration)
ADMM vector per problem, stage interleaved:
        x = (s_0, a_0, s_1, a_1, ..., s_{N-1}, a_{N-1}, s_N),  n = 9 N + 6
blocks  every consecutive 3-vector of x is one prox block: stage k owns blocks
        3k (position), 3k+1 (velocity), 3k+2 (control); the terminal state owns
        3N and 3N+1.  nb = 3 N + 2.
units   nondimensional: time unit 1/n_orbit (mean motion = 1), length unit 1 km.

Python-side array shapes ("math" layout, C order):
        A (Bd,N,6,6)  B (Bd,N,6,3)  c (Bd,N,6)|None  Q (Bd,N+1,6,6)|None
        R (Bd,N,3,3)|None  q (Bq,n)|None  s0 (Bsz,6)
        block_type (nb,) int32   block_par (Bp,nb,8) float64
with Bd, Bq, Bp in {1 (shared), Bsz}.
"""
from __future__ import annotations

import numpy as np

NX = 6
NU = 3

# prox block type codes (mirrored by include/admm_b200.h ADMMB_BLK_*)
BLK_FREE = 0      # g = 0, but the block IS split (x_b = z_b is enforced with g = 0)
BLK_L1 = 1        # g = lam * |v|_1
BLK_L1_BOX = 2    # g = lam * |v|_1 + indicator(lo <= v <= hi)
BLK_L2 = 3        # g = lam * |v|_2          (sum-of-norms fuel)
BLK_L2_BALL = 4   # g = lam * |v|_2 + indicator(|v|_2 <= rad)   (SOC thrust magnitude)
BLK_BOX = 5       # indicator(lo <= v <= hi)
BLK_BALL = 6      # indicator(|v - cen|_2 <= rad)
BLK_POINT = 7     # indicator(v == cen)      (terminal equality as projection)
BLK_NONE = 8      # block takes no part in the splitting (no z, no u): the default for states

# parameter slots inside block_par[..., 8]
PAR_LAM = 0
PAR_RAD = 1
PAR_LO = 2   # lo[3] or cen[3] occupy slots 2..4
PAR_HI = 5   # hi[3] occupy slots 5..7


def n_of(N: int) -> int:
    return 9 * N + 6


def nb_of(N: int) -> int:
    return 3 * N + 2


# --------------------------------------------------------------------------
# dynamics
# --------------------------------------------------------------------------
def cw_continuous(nmm: float = 1.0) -> np.ndarray:
    """Continuous-time Clohessy-Wiltshire system matrix (6x6)."""
    Ac = np.zeros((6, 6))
    Ac[0:3, 3:6] = np.eye(3)
    Ac[3, 0] = 3.0 * nmm * nmm
    Ac[5, 2] = -nmm * nmm
    Ac[3, 4] = 2.0 * nmm
    Ac[4, 3] = -2.0 * nmm
    return Ac


def cw_stm(T: float, nmm: float = 1.0) -> np.ndarray:
    """Closed-form CW state transition matrix Phi(T)."""
    c, s = np.cos(nmm * T), np.sin(nmm * T)
    n = nmm
    P = np.zeros((6, 6))
    P[0:3, 0:3] = [[4 - 3 * c, 0, 0], [6 * (s - n * T), 1, 0], [0, 0, c]]
    P[0:3, 3:6] = [[s / n, 2 * (1 - c) / n, 0],
                   [-2 * (1 - c) / n, (4 * s - 3 * n * T) / n, 0],
                   [0, 0, s / n]]
    P[3:6, 0:3] = [[3 * n * s, 0, 0], [-6 * n * (1 - c), 0, 0], [0, 0, -n * s]]
    P[3:6, 3:6] = [[c, 2 * s, 0], [-2 * s, 4 * c - 3, 0], [0, 0, c]]
    return P


def cw_zoh(T: float, nmm: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
    """(Phi(T), Gamma(T)) for a zero-order-hold acceleration: Gamma = int_0^T Phi(tau) dtau [0;I].

    Closed form (integrating the columns of Phi_rv / Phi_vv)."""
    n = nmm
    c, s = np.cos(n * T), np.sin(n * T)
    Phi = cw_stm(T, nmm)
    G = np.zeros((6, 3))
    # int Phi_rv
    G[0:3, :] = [[(1 - c) / n**2, 2 * (n * T - s) / n**2, 0],
                 [-2 * (n * T - s) / n**2, (4 * (1 - c) - 1.5 * (n * T) ** 2) / n**2, 0],
                 [0, 0, (1 - c) / n**2]]
    # int Phi_vv
    G[3:6, :] = [[s / n, 2 * (1 - c) / n, 0],
                 [-2 * (1 - c) / n, (4 * s - 3 * n * T) / n, 0],
                 [0, 0, s / n]]
    return Phi, G


def elliptic_lvlh_matrix(theta: np.ndarray, e: np.ndarray) -> np.ndarray:
    """Continuous LVLH linearised relative dynamics about a Kepler orbit of eccentricity e at
    true anomaly theta (mean motion 1, mu = 1, a = 1).  Vectorised over the batch: -> (B,6,6)."""
    theta = np.asarray(theta, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    p = 1.0 - e * e
    one_ec = 1.0 + e * np.cos(theta)
    r = p / one_ec
    h = np.sqrt(p)
    w = h / (r * r)                      # theta_dot
    rdot = e * np.sin(theta) / h
    wdot = -2.0 * w * rdot / r
    k = 1.0 / r**3                       # mu / r^3
    Bsz = theta.shape[0]
    Ac = np.zeros((Bsz, 6, 6))
    Ac[:, 0, 3] = Ac[:, 1, 4] = Ac[:, 2, 5] = 1.0
    Ac[:, 3, 0] = w * w + 2.0 * k
    Ac[:, 3, 1] = wdot
    Ac[:, 3, 4] = 2.0 * w
    Ac[:, 4, 0] = -wdot
    Ac[:, 4, 1] = w * w - k
    Ac[:, 4, 3] = -2.0 * w
    Ac[:, 5, 2] = -k
    return Ac


def elliptic_stage_matrices(e: np.ndarray, theta0: np.ndarray, N: int, T: float,
                            substeps: int = 8) -> tuple[np.ndarray, np.ndarray]:
    """RK4-propagated per-problem time-varying (Phi_k, Gamma_k), k = 0..N-1, ZOH input.
    Integrates  Phi' = Ac(theta) Phi,  Gamma' = Ac(theta) Gamma + [0;I],  theta' = w(theta)
    over each stage of length T with `substeps` RK4 steps.  -> A (B,N,6,6), B (B,N,6,3)."""
    e = np.asarray(e, dtype=np.float64)
    theta = np.array(theta0, dtype=np.float64)
    Bsz = e.shape[0]
    A = np.zeros((Bsz, N, 6, 6))
    Bm = np.zeros((Bsz, N, 6, 3))
    Bc = np.zeros((6, 3))
    Bc[3:6, :] = np.eye(3)
    p = 1.0 - e * e
    hh = np.sqrt(p)

    def thdot(th):
        return hh * (1.0 + e * np.cos(th)) ** 2 / (p * p)

    def f(th, Y):
        Ac = elliptic_lvlh_matrix(th, e)
        dY = Ac @ Y
        dY[:, :, 6:9] += Bc
        return dY

    dt = T / substeps
    for k in range(N):
        Y = np.zeros((Bsz, 6, 9))
        Y[:, :, 0:6] = np.eye(6)
        for _ in range(substeps):
            k1t = thdot(theta)
            k1 = f(theta, Y)
            k2t = thdot(theta + 0.5 * dt * k1t)
            k2 = f(theta + 0.5 * dt * k1t, Y + 0.5 * dt * k1)
            k3t = thdot(theta + 0.5 * dt * k2t)
            k3 = f(theta + 0.5 * dt * k2t, Y + 0.5 * dt * k2)
            k4t = thdot(theta + dt * k3t)
            k4 = f(theta + dt * k3t, Y + dt * k3)
            Y = Y + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            theta = theta + (dt / 6.0) * (k1t + 2 * k2t + 2 * k3t + k4t)
        A[:, k] = Y[:, :, 0:6]
        Bm[:, k] = Y[:, :, 6:9]
    return A, Bm


# --------------------------------------------------------------------------
# block descriptors
# --------------------------------------------------------------------------
def make_blocks(N: int, ctrl_type: int, *, lam: float = 0.0, rad: float = 0.0,
                lo: float | np.ndarray = 0.0, hi: float | np.ndarray = 0.0,
                terminal: np.ndarray | None = None,
                state_type: int = BLK_NONE) -> tuple[np.ndarray, np.ndarray]:
    """Shared descriptor table: unsplit (BLK_NONE) states, `ctrl_type` on every control block,
    terminal state pinned to `terminal` (6,) via two POINT blocks (None -> unsplit).
    `state_type=BLK_FREE` reproduces the literal x = z splitting of every entry
    (SURVEY.md section 7.1); it converges about 10x slower and is kept as an option."""
    nb = nb_of(N)
    bt = np.full(nb, state_type, dtype=np.int32)
    bp = np.zeros((1, nb, 8))
    for k in range(N):
        b = 3 * k + 2
        bt[b] = ctrl_type
        bp[0, b, PAR_LAM] = lam
        bp[0, b, PAR_RAD] = rad
        bp[0, b, PAR_LO:PAR_LO + 3] = lo
        bp[0, b, PAR_HI:PAR_HI + 3] = hi
    if terminal is not None:
        terminal = np.asarray(terminal, dtype=np.float64)
        bt[3 * N] = BLK_POINT
        bt[3 * N + 1] = BLK_POINT
        bp[0, 3 * N, PAR_LO:PAR_LO + 3] = terminal[0:3]
        bp[0, 3 * N + 1, PAR_LO:PAR_LO + 3] = terminal[3:6]
    return bt, bp


# --------------------------------------------------------------------------
# the five benchmark configurations
# --------------------------------------------------------------------------
S0_NOMINAL = np.array([1.0, -5.0, 0.5, 0.0, 0.0, 0.0])
S0_SIGMA = np.array([0.1, 0.1, 0.1, 1e-3, 1e-3, 1e-3])

DEFAULT_OPTS = dict(rho=1.0, alpha=1.6, abstol=1e-6, reltol=1e-6, max_iter=5000,
                    adapt_rho=0, adapt_mu=10.0, adapt_tau=2.0, adapt_every=25, adapt_until=0,
                    history=0, xupdate="auto", precision="fp64")


def _dispersed_s0(batch: int, seed: int, spread: float = 1.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return S0_NOMINAL[None, :] + spread * S0_SIGMA[None, :] * rng.standard_normal((batch, 6))


def cfg1_single_impulsive(N: int = 20, dv_max: float = 0.4) -> tuple[dict, dict]:
    """configs[0]: one CW fuel-optimal impulsive rendezvous, L1 cost, box on dv."""
    T = 2.0 * np.pi / N
    Phi = cw_stm(T)
    A = np.broadcast_to(Phi, (1, N, 6, 6)).copy()
    B = np.broadcast_to(Phi[:, 3:6], (1, N, 6, 3)).copy()
    bt, bp = make_blocks(N, BLK_L1_BOX, lam=1.0, lo=-dv_max, hi=dv_max, terminal=np.zeros(6))
    prob = dict(N=N, A=A, B=B, c=None, Q=None, R=None, q=None,
                s0=S0_NOMINAL[None, :].copy(), block_type=bt, block_par=bp)
    opts = dict(DEFAULT_OPTS, rho=1.0, alpha=1.6)
    return prob, opts


def cfg2_cw_batch(batch: int = 4096, N: int = 50, seed: int = 2, dv_max: float = 0.4,
                  spread: float = 1.0) -> tuple[dict, dict]:
    """configs[1]: batch of CW impulsive rendezvous QPs with shared dynamics, dispersed s0."""
    T = 2.0 * np.pi / N
    Phi = cw_stm(T)
    A = np.broadcast_to(Phi, (1, N, 6, 6)).copy()
    B = np.broadcast_to(Phi[:, 3:6], (1, N, 6, 3)).copy()
    bt, bp = make_blocks(N, BLK_L1_BOX, lam=1.0, lo=-dv_max, hi=dv_max, terminal=np.zeros(6))
    prob = dict(N=N, A=A, B=B, c=None, Q=None, R=None, q=None,
                s0=_dispersed_s0(batch, seed, spread), block_type=bt, block_par=bp)
    # rho and alpha tuned once with the oracle (round 2: 1,024 problems, rho in 1 .. 300, alpha in 1.0 .. 1.8) and frozen:
    # 2,048 of 2,048 problems converge to 1e-6 in 2,000 .. 21,200 iterations (median 8,300; of 65,536, one is still running
    # at 40,000); with round 1's rho = 1 one problem in
    # eight was still running at 20,000.  Over-relaxation does not help these LP-like problems.
    opts = dict(DEFAULT_OPTS, rho=40.0, alpha=1.0, max_iter=30000)
    return prob, opts


def cfg3_lowthrust_soc(batch: int = 65536, N: int = 100, seed: int = 3,
                       a_max: float = 2.5) -> tuple[dict, dict]:
    """configs[2]: CW ZOH low-thrust transfers, thrust-magnitude SOC (l2 + ball) on each control."""
    T = 2.0 * np.pi / N
    Phi, Gam = cw_zoh(T)
    A = np.broadcast_to(Phi, (1, N, 6, 6)).copy()
    B = np.broadcast_to(Gam, (1, N, 6, 3)).copy()
    bt, bp = make_blocks(N, BLK_L2_BALL, lam=T, rad=a_max, terminal=np.zeros(6))
    prob = dict(N=N, A=A, B=B, c=None, Q=None, R=None, q=None,
                s0=_dispersed_s0(batch, seed), block_type=bt, block_par=bp)
    # a_max = 2.5: with round 1's 1.5 one dispersed start in ten could not reach the target within the horizon at all
    # (thrust saturated on every stage, primal residual stuck).  rho, alpha tuned once with the oracle: every problem
    # converges in 800 .. 4,000 iterations (median 1,400).
    opts = dict(DEFAULT_OPTS, rho=0.1, alpha=1.6, max_iter=20000)
    return prob, opts


def cfg4_elliptic(batch: int = 16384, N: int = 50, seed: int = 4,
                  a_max: float = 10.0, e_max: float = 0.5) -> tuple[dict, dict]:
    """configs[3]: elliptic-orbit rendezvous with per-problem time-varying STMs (Riccati path).
    e ~ U(0.05, 0.5), a_max = 10: with round 1's e up to 0.7 and a_max = 3 a quarter of the problems were infeasible
    (thrust saturated on every stage).  rho, alpha tuned once with the oracle: all of 256 converge (median 1,700 iterations, 99 % below 6,500); of 16,384, 99.6 % by 15,000 and all but 2 by 40,000."""
    rng = np.random.Generator(np.random.PCG64(seed))
    e = rng.uniform(0.05, e_max, batch)
    th0 = rng.uniform(0.0, 2.0 * np.pi, batch)
    T = 2.0 * np.pi / N
    A, B = elliptic_stage_matrices(e, th0, N, T)
    bt, bp = make_blocks(N, BLK_L2_BALL, lam=T, rad=a_max, terminal=np.zeros(6))
    s0 = S0_NOMINAL[None, :] + S0_SIGMA[None, :] * rng.standard_normal((batch, 6))
    prob = dict(N=N, A=A, B=B, c=None, Q=None, R=None, q=None,
                s0=s0, block_type=bt, block_par=bp, meta=dict(e=e, theta0=th0))
    opts = dict(DEFAULT_OPTS, rho=0.05, alpha=1.6, max_iter=30000)
    return prob, opts


def cfg5_montecarlo(batch: int = 1048576, N: int = 50, seed: int = 5,
                    dv_max: float = 1.0, spread: float = 5.0) -> tuple[dict, dict]:
    """configs[4]: Monte Carlo dispersion sweep (5x the dispersion of config 2, so iteration counts vary by more than
    10x), adaptive rho + per-problem early exit.  Residual balancing with the textbook mu = 10 drives rho of these
    LP-like problems to ~1, where they converge 2-3x slower than at rho = 40 (measured with the oracle, round 2), so
    the trigger only corrects gross imbalance (mu = 1000) during the first 1000 iterations."""
    prob, opts = cfg2_cw_batch(batch, N, seed, dv_max, spread=spread)
    opts = dict(opts, adapt_rho=1, adapt_mu=1000.0, adapt_tau=2.0, adapt_every=100, adapt_until=1000, max_iter=60000)
    return prob, opts


def lqr_tracking(batch: int = 64, N: int = 30, seed: int = 7, per_problem: bool = False,
                 with_affine: bool = True) -> tuple[dict, dict]:
    """Extra coverage case (not a benchmark config): quadratic cost (Q, R != 0), linear cost q,
    affine dynamics term c, box-constrained controls and a terminal ball.  Exercises the
    rho-dependent factor, the q/rho term and the c_k path."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T = 2.0 * np.pi / N
    Phi, Gam = cw_zoh(T)
    Bd = batch if per_problem else 1
    A = np.broadcast_to(Phi, (Bd, N, 6, 6)).copy()
    B = np.broadcast_to(Gam, (Bd, N, 6, 3)).copy()
    if per_problem:
        A += 1e-2 * rng.standard_normal(A.shape)
        B += 1e-2 * rng.standard_normal(B.shape)
    c = 1e-2 * rng.standard_normal((Bd, N, 6)) if with_affine else None
    Q = np.broadcast_to(np.diag([1e-2] * 3 + [1e-3] * 3), (Bd, N + 1, 6, 6)).copy()
    R = np.broadcast_to(np.eye(3), (Bd, N, 3, 3)).copy()
    n = n_of(N)
    q = 1e-2 * rng.standard_normal((batch, n))
    bt, bp = make_blocks(N, BLK_BOX, lo=-0.5, hi=0.5)
    bt[3 * N] = BLK_BALL
    bp[0, 3 * N, PAR_RAD] = 0.05
    bt[3 * N + 1] = BLK_BALL
    bp[0, 3 * N + 1, PAR_RAD] = 0.01
    prob = dict(N=N, A=A, B=B, c=c, Q=Q, R=R, q=q, s0=0.05 * _dispersed_s0(batch, seed),
                block_type=bt, block_par=bp)
    opts = dict(DEFAULT_OPTS, rho=1.0, alpha=1.5)
    return prob, opts


def scp_nonlinear_rendezvous(batch: int = 4096, N: int = 50, seed: int = 6, scale: float = 20.0,
                             R0: float = 7000.0, substeps: int = 4, max_pass: int = 20) -> tuple[dict, dict, dict]:
    """SURVEY 8(f-4) workload: low-thrust rendezvous from tens to hundreds of km (config 3's dispersed starts, 3 sigma,
    times `scale`) about a chief on a circular orbit of radius R0 km -- far enough for the Clohessy-Wiltshire model to be
    visibly wrong, so the convex subproblem is re-linearised about each problem's own trajectory (sequential convex
    programming).  Thrust-magnitude SOC on every control, terminal point, one orbit.  -> (prob, scp, opts); the stage
    matrices are produced by the solver (Solver.scp_solve / oracle.scp_ocp.scp_solve), prob carries N instead.
    Parameters tuned once with the oracle (1,024 problems, N = 50) and frozen: every trajectory converges, in 4 .. 13 passes
    (median 6) and 2 k .. 23 k ADMM iterations in total (median 6 k); thrust bound 3.5 x scale (with config 3's 2.5 one start
    in a thousand saturates the thrust on every stage and its subproblems never converge).  The
    trajectory tolerance sits above what the ADMM tolerance (1e-6 relative on 2-norms of ~1e3) leaves undetermined in
    x: with tol_rel = 1e-6 one problem in a thousand re-solved its already-converged subproblem in one iteration per
    pass and moved by 1.1e-4 each time, just above the threshold, until max_pass."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T = 2.0 * np.pi / N
    s0 = scale * (S0_NOMINAL[None, :] + 3.0 * S0_SIGMA[None, :] * rng.standard_normal((batch, 6)))
    bt, bp = make_blocks(N, BLK_L2_BALL, lam=T, rad=3.5 * scale, terminal=np.zeros(6))
    prob = dict(N=N, A=None, B=None, c=None, Q=None, R=None, q=None, s0=s0, block_type=bt, block_par=bp)
    scp = dict(model="nl_circular", T=T, R0=R0, nmm=1.0, substeps=substeps, max_pass=max_pass, tol_abs=1e-5, tol_rel=1e-5)
    # rho: config 3's 0.1 carried over by the scaling of the problem would be 0.1 / scale; 0.01 measured best with the
    # oracle (96 problems: slowest problem 57 k ADMM iterations in total against 118 k at rho = 0.1).  max_iter = 2,000 caps
    # the convex solves of the early passes, whose subproblems are far from the final one (inexact SCP): same passes
    # (5 .. 8 on those 96), same nonlinear terminal miss, slowest problem 13 k iterations in total.
    opts = dict(DEFAULT_OPTS, rho=0.01 * 20.0 / scale, alpha=1.6, max_iter=2000)
    return prob, scp, opts


def scp_nonlinear_impulsive(batch: int = 1024, N: int = 30, seed: int = 7, scale: float = 20.0, R0: float = 7000.0,
                            substeps: int = 4, max_pass: int = 40, dv_max: float = 0.6) -> tuple[dict, dict, dict]:
    """SURVEY 8(f-4), the impulsive family (configs 1, 2, 5 with nonlinear relative dynamics): L1 fuel cost + box on each
    velocity increment, applied at the start of a stage and followed by a coast; terminal point; starts as in
    scp_nonlinear_rendezvous.  These LP-like subproblems need several thousand ADMM iterations each (as config 2 does), so
    max_iter caps a pass at 3,000 and the trajectory takes 10 .. 30 passes.  Tuned once with the oracle (64 problems,
    N = 50: all converge with rho = 0.5 at scale 20; rho = 0.1 .. 1 with a 2,000 cap leaves one in ten unconverged)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T = 2.0 * np.pi / N
    s0 = scale * (S0_NOMINAL[None, :] + 3.0 * S0_SIGMA[None, :] * rng.standard_normal((batch, 6)))
    bt, bp = make_blocks(N, BLK_L1_BOX, lam=1.0, lo=-dv_max * scale, hi=dv_max * scale, terminal=np.zeros(6))
    prob = dict(N=N, A=None, B=None, c=None, Q=None, R=None, q=None, s0=s0, block_type=bt, block_par=bp)
    scp = dict(model="nl_circular", control="impulsive", T=T, R0=R0, nmm=1.0, substeps=substeps, max_pass=max_pass,
               tol_abs=1e-5, tol_rel=1e-5)
    opts = dict(DEFAULT_OPTS, rho=0.5 * 20.0 / scale, alpha=1.0, max_iter=3000)
    return prob, scp, opts


def scp_nonlinear_elliptic(batch: int = 1024, N: int = 50, seed: int = 8, scale: float = 20.0, R0: float = 7000.0,
                           substeps: int = 4, max_pass: int = 30, e_max: float = 0.5) -> tuple[dict, dict, dict]:
    """SURVEY 8(f-4), config 4's family with nonlinear relative dynamics: the chief is on a Kepler orbit (e ~ U(0.05, e_max),
    theta0 ~ U(0, 2 pi), semi-major axis R0 km), thrust-magnitude SOC on every control, terminal point, one orbit.  Thrust
    bound carried over from config 4 by the scaling of the problem (a_max = 10 x scale); rho checked once with the oracle (32
    problems, N = 30 and 50: all converge in 4 .. 20 passes; 0.003 needs the fewest ADMM iterations of 0.003 / 0.01 / 0.03)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T = 2.0 * np.pi / N
    e = rng.uniform(0.05, e_max, batch)
    th0 = rng.uniform(0.0, 2.0 * np.pi, batch)
    s0 = scale * (S0_NOMINAL[None, :] + 3.0 * S0_SIGMA[None, :] * rng.standard_normal((batch, 6)))
    bt, bp = make_blocks(N, BLK_L2_BALL, lam=T, rad=10.0 * scale, terminal=np.zeros(6))
    prob = dict(N=N, A=None, B=None, c=None, Q=None, R=None, q=None, s0=s0, block_type=bt, block_par=bp)
    scp = dict(model="nl_elliptic", control="zoh", T=T, R0=R0, nmm=1.0, substeps=substeps, max_pass=max_pass,
               tol_abs=1e-5, tol_rel=1e-5, e=e, theta0=th0)
    opts = dict(DEFAULT_OPTS, rho=0.003 * 20.0 / scale, alpha=1.6, max_iter=2000)
    return prob, scp, opts


def batch_size(prob: dict) -> int:
    return int(prob["s0"].shape[0])
