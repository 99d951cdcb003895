"""ORACLE (test infrastructure) -- sequential convex programming (SCP) outer loop around the batched ADMM oracle.

SURVEY.md 8(f-4): "SCP outer loop (re-linearise nonlinear dynamics, call batched ADMM per pass)".  The CUDA library
runs this loop on the resident batch (`admmb_scp_solve`, csrc/scp.cuh); this file is the CPU statement of exactly the
same arithmetic, so that the device result can be compared BIT FOR BIT, pass by pass:

  * the nonlinear model: relative motion of a deputy about a chief on a circular orbit of radius R0 (LVLH frame, x
    radial, y along-track, z cross-track, mean motion n, mu = n^2 R0^3), full two-body gravity, zero-order-hold
    thrust acceleration a (scp["control"] = "zoh") or a velocity increment at the start of each stage followed by a
    coast ("impulsive": the fuel-optimal impulsive rendezvous of configs 1, 2, 5).  Its linearisation at r = 0 is the
    Clohessy-Wiltshire model of configs 1-3.  scp["model"] = "nl_elliptic": the chief is on a Kepler orbit of eccentricity
    e[p] starting at true anomaly theta0[p] (semi-major axis R0, time unit 1 / mean motion; the LVLH frame's rotation rate,
    its derivative and mu / R^3 follow theta(t), integrated by the same RK4) -- linearised at r = 0: config 4's model;
  * per pass, per stage: classical RK4 (`substeps` per stage) of the state together with its variational equations
    (one column of [Phi | Gamma] at a time, as the device does) about the reference (s_ref_k, a_ref_k):
        A_k = dF/ds, B_k = dF/da, c_k = F(s_ref_k, a_ref_k) - A_k s_ref_k - B_k a_ref_k;
  * the convex subproblem (same cost / constraint blocks, per-problem affine time-varying dynamics) goes to the ADMM
    oracle (oracle/admm_ocp_cpu.c through oracle/cpu.py), warm-started from the previous pass's (z, u);
  * a problem leaves the loop when max|x - x_ref| <= tol_abs + tol_rel max|x| AND its convex solve converged (or after
    max_pass passes) -- so opts.max_iter may cap the early passes, whose subproblems are far from the final one
    (inexact SCP: 3-9x fewer ADMM iterations in total); the first reference is the free drift from s0 (a = 0),
    propagated with the same RK4.

Every value is produced by IEEE-754 double multiplications, additions, divisions and square roots in the order
written here (NumPy element-wise operations never fuse; the CUDA side is compiled with -fmad=false).  The gravity
difference n^2 (1 - R0^3 / d^3) is evaluated in the cancellation-free form q (3 + 3q + q^2) / (w^1.5 (1 + w^1.5)).

The reference (/root/reference/README.md:1-2) has no SCP loop (it has no code): parity unpinned.  What pins this
file: tests/test_oracle.py checks the stage map against scipy solve_ivp of the same nonlinear equations, A_k / B_k
against finite differences, the r -> 0 limit against the CW closed forms, and that the converged SCP controls, flown
through the nonlinear dynamics, reach the target.

Only tests/, __graft_entry__.smoke() and bench.py's checker may import this module."""
from __future__ import annotations

import numpy as np

MODEL_NL_CIRCULAR = 1


def _coeffs(s, R0, n2):
    """Gravity-gradient coefficients at the state s (list of six (B,) arrays)."""
    rx = R0 + s[0]
    q = ((((2.0 * R0) * s[0] + s[0] * s[0]) + s[1] * s[1]) + s[2] * s[2]) / (R0 * R0)
    w = 1.0 + q
    w32 = w * np.sqrt(w)
    k = n2 / w32                                               # mu / d^3
    g = (n2 * (q * ((3.0 + 3.0 * q) + q * q))) / (w32 * (1.0 + w32))   # n^2 - mu / d^3
    m = (3.0 * k) / ((R0 * R0) * w)                            # 3 mu / d^5
    return rx, k, g, m


def _f_state(cf, s, a, tn):
    rx, k, g, _ = cf
    return [s[3], s[4], s[5],
            (tn * s[4] + g * rx) + a[0],
            ((-tn) * s[3] + g * s[1]) + a[1],
            (-k) * s[2] + a[2]]


def _jac(cf, s):
    rx, k, g, m = cf
    return (g + m * (rx * rx), m * (rx * s[1]), m * (rx * s[2]), g + m * (s[1] * s[1]), m * (s[1] * s[2]),
            (-k) + m * (s[2] * s[2]))


def _f_col(J, y, tn, forced_row):
    j30, j31, j32, j41, j42, j52 = J
    d3 = ((j30 * y[0] + j31 * y[1]) + j32 * y[2]) + tn * y[4]
    d4 = ((j31 * y[0] + j41 * y[1]) + j42 * y[2]) + (-tn) * y[3]
    d5 = (j32 * y[0] + j42 * y[1]) + j52 * y[2]
    dy = [y[3], y[4], y[5], d3, d4, d5]
    if forced_row >= 0:
        dy[forced_row] = dy[forced_row] + 1.0
    return dy


def linearise_stage(s_ref, a_ref, T, substeps, nmm, R0, impulsive=False):
    """One stage about (s_ref [B,6], a_ref [B,3]) -> F [B,6], A [B,6,6], Bm [B,6,3], c [B,6].
    impulsive: the control is a velocity increment applied at the start of the stage, followed by a coast."""
    if impulsive:
        return _linearise_stage_impulsive(s_ref, a_ref, T, substeps, nmm, R0)
    Bsz = s_ref.shape[0]
    n = np.float64(nmm)
    R0 = np.float64(R0)
    n2 = n * n
    tn = 2.0 * n
    dt = np.float64(T) / np.float64(substeps)
    hdt = 0.5 * dt
    dt6 = dt / 6.0
    s = [s_ref[:, i].copy() for i in range(6)]
    a = [a_ref[:, i].copy() for i in range(3)]
    cols = []
    for j in range(9):
        y = [np.zeros(Bsz) for _ in range(6)]
        if j < 6:
            y[j] = np.ones(Bsz)
        cols.append(y)
    for _ in range(substeps):
        c1 = _coeffs(s, R0, n2)
        k1 = _f_state(c1, s, a, tn)
        s2 = [s[i] + hdt * k1[i] for i in range(6)]
        c2 = _coeffs(s2, R0, n2)
        k2 = _f_state(c2, s2, a, tn)
        s3 = [s[i] + hdt * k2[i] for i in range(6)]
        c3 = _coeffs(s3, R0, n2)
        k3 = _f_state(c3, s3, a, tn)
        s4 = [s[i] + dt * k3[i] for i in range(6)]
        c4 = _coeffs(s4, R0, n2)
        k4 = _f_state(c4, s4, a, tn)
        J1, J2, J3, J4 = _jac(c1, s), _jac(c2, s2), _jac(c3, s3), _jac(c4, s4)
        for j in range(9):
            fr = j - 3 if j >= 6 else -1
            y = cols[j]
            l1 = _f_col(J1, y, tn, fr)
            l2 = _f_col(J2, [y[i] + hdt * l1[i] for i in range(6)], tn, fr)
            l3 = _f_col(J3, [y[i] + hdt * l2[i] for i in range(6)], tn, fr)
            l4 = _f_col(J4, [y[i] + dt * l3[i] for i in range(6)], tn, fr)
            cols[j] = [y[i] + dt6 * (((l1[i] + 2.0 * l2[i]) + 2.0 * l3[i]) + l4[i]) for i in range(6)]
        s = [s[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]) for i in range(6)]
    F = np.stack(s, axis=1)
    A = np.zeros((Bsz, 6, 6))
    Bm = np.zeros((Bsz, 6, 3))
    c = [s[i].copy() for i in range(6)]
    for j in range(9):
        v = s_ref[:, j] if j < 6 else a_ref[:, j - 6]
        for i in range(6):
            if j < 6:
                A[:, i, j] = cols[j][i]
            else:
                Bm[:, i, j - 6] = cols[j][i]
            c[i] = c[i] - cols[j][i] * v
    return F, A, Bm, np.stack(c, axis=1)


def _linearise_stage_impulsive(s_ref, a_ref, T, substeps, nmm, R0):
    """s+ = F(s + [0; dv]) with F the coast over T: A = dF/ds at the post-impulse state, B = A[:, 3:6],
    c = F - A s_ref - B dv_ref, subtracted column by column (state entry j, then -- for the velocity columns -- impulse
    entry j - 3), as the device does."""
    Bsz = s_ref.shape[0]
    n = np.float64(nmm)
    R0 = np.float64(R0)
    n2 = n * n
    tn = 2.0 * n
    dt = np.float64(T) / np.float64(substeps)
    hdt = 0.5 * dt
    dt6 = dt / 6.0
    s = [s_ref[:, i].copy() for i in range(3)] + [s_ref[:, 3 + i] + a_ref[:, i] for i in range(3)]
    a = [np.zeros(Bsz) for _ in range(3)]
    cols = []
    for j in range(6):
        y = [np.zeros(Bsz) for _ in range(6)]
        y[j] = np.ones(Bsz)
        cols.append(y)
    for _ in range(substeps):
        c1 = _coeffs(s, R0, n2)
        k1 = _f_state(c1, s, a, tn)
        s2 = [s[i] + hdt * k1[i] for i in range(6)]
        c2 = _coeffs(s2, R0, n2)
        k2 = _f_state(c2, s2, a, tn)
        s3 = [s[i] + hdt * k2[i] for i in range(6)]
        c3 = _coeffs(s3, R0, n2)
        k3 = _f_state(c3, s3, a, tn)
        s4 = [s[i] + dt * k3[i] for i in range(6)]
        c4 = _coeffs(s4, R0, n2)
        k4 = _f_state(c4, s4, a, tn)
        J1, J2, J3, J4 = _jac(c1, s), _jac(c2, s2), _jac(c3, s3), _jac(c4, s4)
        for j in range(6):
            y = cols[j]
            l1 = _f_col(J1, y, tn, -1)
            l2 = _f_col(J2, [y[i] + hdt * l1[i] for i in range(6)], tn, -1)
            l3 = _f_col(J3, [y[i] + hdt * l2[i] for i in range(6)], tn, -1)
            l4 = _f_col(J4, [y[i] + dt * l3[i] for i in range(6)], tn, -1)
            cols[j] = [y[i] + dt6 * (((l1[i] + 2.0 * l2[i]) + 2.0 * l3[i]) + l4[i]) for i in range(6)]
        s = [s[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]) for i in range(6)]
    F = np.stack(s, axis=1)
    A = np.zeros((Bsz, 6, 6))
    Bm = np.zeros((Bsz, 6, 3))
    c = [s[i].copy() for i in range(6)]
    for j in range(6):
        for i in range(6):
            A[:, i, j] = cols[j][i]
            c[i] = c[i] - cols[j][i] * s_ref[:, j]
        if j >= 3:
            for i in range(6):
                Bm[:, i, j - 3] = cols[j][i]
                c[i] = c[i] - cols[j][i] * a_ref[:, j - 3]
    return F, A, Bm, np.stack(c, axis=1)


# ---- model "nl_elliptic": chief on a Kepler orbit of eccentricity e (semi-major axis R0, time unit 1 / mean motion) ----
def _ell_coef(th, e, p, h, a):
    """Orbit quantities at true anomaly th (the expressions of oracle/gen_ocp.py _elliptic_coeffs): rotation rate w and its
    derivative wd of the LVLH frame, K = mu / R^3, chief radius R = a r."""
    from .gen_ocp import det_sincos
    st, ct = det_sincos(th)
    one_ec = 1.0 + e * ct
    r = p / one_ec
    w = h / (r * r)
    rdot = (e * st) / h
    wd = ((-2.0 * w) * rdot) / r
    K = 1.0 / ((r * r) * r)
    R = a * r
    return w, wd, K, R


def _ell_grav(oc, s):
    """State-dependent part: K (1 - (R/d)^3) and 3 K / (R^2 w^2.5), cancellation-free as in _coeffs."""
    w, wd, K, R = oc
    rx = R + s[0]
    q = ((((2.0 * R) * s[0] + s[0] * s[0]) + s[1] * s[1]) + s[2] * s[2]) / (R * R)
    w_ = 1.0 + q
    w32 = w_ * np.sqrt(w_)
    gf = (q * ((3.0 + 3.0 * q) + q * q)) / (w32 * (1.0 + w32))
    Kg = K * gf
    m = (3.0 * K) / (((R * R) * w_) * w32)
    return rx, Kg, m


def _ell_f_state(oc, gc, s, a):
    w, wd, K, _ = oc
    rx, Kg, _ = gc
    tw = 2.0 * w
    wK = w * w - K
    return [s[3], s[4], s[5],
            ((((tw * s[4]) + (wd * s[1])) + (wK * s[0])) + (Kg * rx)) + a[0],
            (((((-tw) * s[3]) + ((-wd) * s[0])) + (wK * s[1])) + (Kg * s[1])) + a[1],
            (((-K) * s[2]) + (Kg * s[2])) + a[2]]


def _ell_f_col(oc, gc, s, y, forced_row):
    w, wd, K, _ = oc
    rx, Kg, m = gc
    tw = 2.0 * w
    wK = w * w - K
    j30 = (wK + Kg) + m * (rx * rx); j31 = wd + m * (rx * s[1]); j32 = m * (rx * s[2])
    j40 = (-wd) + m * (rx * s[1]); j41 = (wK + Kg) + m * (s[1] * s[1]); j42 = m * (s[1] * s[2])
    j50 = m * (rx * s[2]); j51 = m * (s[1] * s[2]); j52 = ((-K) + Kg) + m * (s[2] * s[2])
    dy = [y[3], y[4], y[5],
          (((j30 * y[0]) + (j31 * y[1])) + (j32 * y[2])) + tw * y[4],
          (((j40 * y[0]) + (j41 * y[1])) + (j42 * y[2])) + (-tw) * y[3],
          ((j50 * y[0]) + (j51 * y[1])) + (j52 * y[2])]
    if forced_row >= 0:
        dy[forced_row] = dy[forced_row] + 1.0
    return dy


def theta_table(e, theta0, N, T, substeps):
    """True anomaly at the start of every stage [B, N] (+ the final one): RK4 of theta' = w(theta), the orbit's own clock."""
    e = np.asarray(e, dtype=np.float64)
    th = np.array(theta0, dtype=np.float64)
    p = 1.0 - e * e
    h = np.sqrt(p)
    dt = np.float64(T) / np.float64(substeps)
    hdt = 0.5 * dt
    dt6 = dt / 6.0
    tab = np.zeros((e.shape[0], N))
    one = np.float64(1.0)
    for k in range(N):
        tab[:, k] = th
        for _ in range(substeps):
            w1 = _ell_coef(th, e, p, h, one)[0]
            w2 = _ell_coef(th + hdt * w1, e, p, h, one)[0]
            w3 = _ell_coef(th + hdt * w2, e, p, h, one)[0]
            w4 = _ell_coef(th + dt * w3, e, p, h, one)[0]
            th = th + dt6 * (((w1 + 2.0 * w2) + 2.0 * w3) + w4)
    return tab


def linearise_stage_elliptic(s_ref, a_ref, th_k, e, T, substeps, R0, impulsive=False):
    """linearise_stage for the chief on an elliptic orbit; th_k [B]: true anomaly at the start of the stage."""
    Bsz = s_ref.shape[0]
    e = np.asarray(e, dtype=np.float64)
    a_sma = np.float64(R0)
    p = 1.0 - e * e
    h = np.sqrt(p)
    dt = np.float64(T) / np.float64(substeps)
    hdt = 0.5 * dt
    dt6 = dt / 6.0
    th = np.array(th_k, dtype=np.float64)
    if impulsive:
        s = [s_ref[:, i].copy() for i in range(3)] + [s_ref[:, 3 + i] + a_ref[:, i] for i in range(3)]
        a = [np.zeros(Bsz) for _ in range(3)]
        ncol = 6
    else:
        s = [s_ref[:, i].copy() for i in range(6)]
        a = [a_ref[:, i].copy() for i in range(3)]
        ncol = 9
    cols = []
    for j in range(ncol):
        y = [np.zeros(Bsz) for _ in range(6)]
        if j < 6:
            y[j] = np.ones(Bsz)
        cols.append(y)
    for _ in range(substeps):
        o1 = _ell_coef(th, e, p, h, a_sma)
        o2 = _ell_coef(th + hdt * o1[0], e, p, h, a_sma)
        o3 = _ell_coef(th + hdt * o2[0], e, p, h, a_sma)
        o4 = _ell_coef(th + dt * o3[0], e, p, h, a_sma)
        g1 = _ell_grav(o1, s)
        k1 = _ell_f_state(o1, g1, s, a)
        s2 = [s[i] + hdt * k1[i] for i in range(6)]
        g2 = _ell_grav(o2, s2)
        k2 = _ell_f_state(o2, g2, s2, a)
        s3 = [s[i] + hdt * k2[i] for i in range(6)]
        g3 = _ell_grav(o3, s3)
        k3 = _ell_f_state(o3, g3, s3, a)
        s4 = [s[i] + dt * k3[i] for i in range(6)]
        g4 = _ell_grav(o4, s4)
        k4 = _ell_f_state(o4, g4, s4, a)
        for j in range(ncol):
            fr = j - 3 if j >= 6 else -1
            y = cols[j]
            l1 = _ell_f_col(o1, g1, s, y, fr)
            l2 = _ell_f_col(o2, g2, s2, [y[i] + hdt * l1[i] for i in range(6)], fr)
            l3 = _ell_f_col(o3, g3, s3, [y[i] + hdt * l2[i] for i in range(6)], fr)
            l4 = _ell_f_col(o4, g4, s4, [y[i] + dt * l3[i] for i in range(6)], fr)
            cols[j] = [y[i] + dt6 * (((l1[i] + 2.0 * l2[i]) + 2.0 * l3[i]) + l4[i]) for i in range(6)]
        s = [s[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]) for i in range(6)]
        th = th + dt6 * (((o1[0] + 2.0 * o2[0]) + 2.0 * o3[0]) + o4[0])
    F = np.stack(s, axis=1)
    A = np.zeros((Bsz, 6, 6))
    Bm = np.zeros((Bsz, 6, 3))
    c = [s[i].copy() for i in range(6)]
    for j in range(ncol):
        v = s_ref[:, j] if j < 6 else a_ref[:, j - 6]
        for i in range(6):
            if j < 6:
                A[:, i, j] = cols[j][i]
            else:
                Bm[:, i, j - 6] = cols[j][i]
            c[i] = c[i] - cols[j][i] * v
        if impulsive and j >= 3:
            for i in range(6):
                Bm[:, i, j - 3] = cols[j][i]
                c[i] = c[i] - cols[j][i] * a_ref[:, j - 3]
    return F, A, Bm, np.stack(c, axis=1)


def _stage(scp, s_ref, a_ref, k, idx=None):
    """One stage of the model scp["model"] ("nl_circular" | "nl_elliptic"); idx: the problems s_ref / a_ref belong to."""
    if scp.get("model", "nl_circular") == "nl_elliptic":
        tab = scp["_theta_tab"]
        e = np.asarray(scp["e"], dtype=np.float64)
        if idx is not None:
            tab, e = tab[idx], e[idx]
        return linearise_stage_elliptic(s_ref, a_ref, tab[:, k], e, scp["T"], scp.get("substeps", 8), scp["R0"],
                                        _impulsive(scp))
    return linearise_stage(s_ref, a_ref, scp["T"], scp.get("substeps", 8), scp.get("nmm", 1.0), scp["R0"], _impulsive(scp))


def with_theta_table(scp: dict, N: int) -> dict:
    """scp + the true anomaly at the start of every stage (model "nl_elliptic" only)."""
    if scp.get("model", "nl_circular") == "nl_elliptic" and "_theta_tab" not in scp:
        scp = dict(scp, _theta_tab=theta_table(scp["e"], scp["theta0"], N, scp["T"], scp.get("substeps", 8)))
    return scp


def _impulsive(scp) -> bool:
    return scp.get("control", "zoh") == "impulsive"


def linearise(xref, N, scp, idx=None):
    """Every stage about the reference trajectory xref [B, 9N+6] -> A [B,N,6,6], Bm [B,N,6,3], c [B,N,6].
    idx: which problems of the batch the rows of xref are (per-problem model parameters, "nl_elliptic")."""
    Bsz = xref.shape[0]
    A = np.zeros((Bsz, N, 6, 6))
    Bm = np.zeros((Bsz, N, 6, 3))
    c = np.zeros((Bsz, N, 6))
    scp = with_theta_table(scp, N)
    for k in range(N):
        _, A[:, k], Bm[:, k], c[:, k] = _stage(scp, xref[:, 9 * k:9 * k + 6], xref[:, 9 * k + 6:9 * k + 9], k, idx)
    return A, Bm, c


def shoot(s0, controls, N, scp):
    """Nonlinear trajectory from s0 [B,6] under `controls` [B,N,3] (None: free drift), linearised on the way.
    -> xref [B, 9N+6], A, Bm, c."""
    Bsz = s0.shape[0]
    xref = np.zeros((Bsz, 9 * N + 6))
    A = np.zeros((Bsz, N, 6, 6))
    Bm = np.zeros((Bsz, N, 6, 3))
    c = np.zeros((Bsz, N, 6))
    s = np.array(s0, dtype=np.float64)
    scp = with_theta_table(scp, N)
    for k in range(N):
        a = np.zeros((Bsz, 3)) if controls is None else controls[:, k]
        xref[:, 9 * k:9 * k + 6] = s
        xref[:, 9 * k + 6:9 * k + 9] = a
        s, A[:, k], Bm[:, k], c[:, k] = _stage(scp, s, a, k)
    xref[:, 9 * N:] = s
    return xref, A, Bm, c


def scp_solve(prob: dict, scp: dict, opts: dict, solve=None):
    """SCP on a batch.  prob: N, s0 [B,6], block_type, block_par, optional q / Q / R as for the ADMM oracle (A, B, c
    are produced here).  scp: dict(T, R0, nmm=1, substeps=8, max_pass, tol_abs, tol_rel, control="zoh" | "impulsive").
    -> x, z, u [B,n], info dict(passes, scp_status (0 converged, 1 max_pass), step, iters_total, iters, status,
    hist_step [B,max_pass] (NaN after a problem's exit))."""
    if solve is None:
        from . import cpu
        solve = cpu.solve
    N = int(prob["N"])
    s0 = np.asarray(prob["s0"], dtype=np.float64)
    Bsz = s0.shape[0]
    n = 9 * N + 6
    max_pass = int(scp["max_pass"])
    tol_abs, tol_rel = float(scp.get("tol_abs", 0.0)), float(scp.get("tol_rel", 0.0))
    scp = with_theta_table(scp, N)
    xref, A, Bm, c = shoot(s0, None, N, scp)
    x = np.zeros((Bsz, n)); z = np.zeros((Bsz, n)); u = np.zeros((Bsz, n))
    passes = np.zeros(Bsz, dtype=np.int32)
    scp_status = np.ones(Bsz, dtype=np.int32)
    step_out = np.full(Bsz, np.nan)
    iters_total = np.zeros(Bsz, dtype=np.int64)
    iters = np.zeros(Bsz, dtype=np.int32)
    status = np.zeros(Bsz, dtype=np.int32)
    hist_step = np.full((Bsz, max_pass), np.nan)
    active = np.ones(Bsz, dtype=bool)
    sub_of = lambda a, idx: a if (a is None or a.shape[0] == 1) else a[idx]   # noqa: E731
    for p in range(1, max_pass + 1):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        if p > 1:
            A[idx], Bm[idx], c[idx] = linearise(xref[idx], N, scp, idx)
        sub = dict(N=N, A=A[idx], B=Bm[idx], c=c[idx], Q=sub_of(prob.get("Q"), idx), R=sub_of(prob.get("R"), idx),
                   q=sub_of(prob.get("q"), idx), s0=s0[idx], block_type=prob["block_type"],
                   block_par=sub_of(prob["block_par"], idx))
        if p > 1:
            sub["z0"], sub["u0"] = z[idx], u[idx]
        else:                                         # an initial warm start given with the problem serves the first pass
            for key in ("z0", "u0"):
                if prob.get(key) is not None:
                    sub[key] = np.asarray(prob[key], dtype=np.float64)[idx]
        xs, zs, us, h = solve(sub, opts)
        x[idx], z[idx], u[idx] = xs, zs, us
        iters[idx], status[idx] = h["iters"], h["status"]
        iters_total[idx] += h["iters"]
        step = np.abs(xs - xref[idx]).max(axis=1)
        scale = np.abs(xs).max(axis=1)
        xref[idx] = xs
        passes[idx] = p
        step_out[idx] = step
        hist_step[idx, p - 1] = step
        done = (step <= tol_abs + tol_rel * scale) & (h["status"] == 0)    # the last convex solve must itself have converged
        scp_status[idx[done]] = 0
        active[idx[done]] = False
    return x, z, u, dict(passes=passes, scp_status=scp_status, step=step_out, iters_total=iters_total, iters=iters,
                         status=status, hist_step=hist_step, xref=xref)


def propagate_nonlinear(s0, controls, N, scp):
    """Terminal state of the nonlinear dynamics under the given controls (tests)."""
    xref, _, _, _ = shoot(s0, controls, N, scp)
    return xref[:, 9 * N:]
