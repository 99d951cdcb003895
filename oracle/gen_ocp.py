"""ORACLE (test infrastructure) -- stage-matrix generators in a spelled-out operation order.

SURVEY.md 8(f-1): the CUDA library generates (A_k, B_k) on the device from a few parameters per problem
(`admmb_upload_generated`, csrc/generators.cuh) instead of taking 21.6 KB per problem over PCIe.  This file is the
CPU statement of exactly the arithmetic those kernels perform: every value is produced by IEEE-754 double
multiplications, additions, divisions, square roots and rint() in the order written here (NumPy element-wise
operations never fuse a multiply with an add; the CUDA side is compiled with -fmad=false), and the sine / cosine
are the polynomials below instead of a libm call, so the device output can be compared BIT FOR BIT.

The reference (/root/reference/README.md:1-2) has no generator (it has no code): parity unpinned.  What pins these
functions is `admm-library_b200/problems.py` (closed forms / RK4 written with NumPy matmul and libm trig, the
generators the round-1/2 workloads were built with): tests/test_oracle.py holds the two within 1e-13, and
problems.py itself is checked there against scipy expm / solve_ivp.

Only tests/, __graft_entry__.smoke() and bench.py's checker may import this module."""
from __future__ import annotations

import numpy as np

# ---- sine / cosine: Cody-Waite reduction by pi/2 in three parts + Horner polynomials on [-pi/4, pi/4] --------------
# (coefficients of the fdlibm k_sin / k_cos kernels; evaluated here with separate multiplies and adds)
PIO2_1 = 1.57079632673412561417e+00   # first 33 bits of pi/2
PIO2_2 = 6.07710050630396597660e-11   # next 33 bits
PIO2_3 = 2.02226624871116645580e-21   # next 33 bits
PIO2_3T = 8.47842766036889956997e-32  # pi/2 - (PIO2_1 + PIO2_2 + PIO2_3)
TWO_OVER_PI = 6.36619772367581382433e-01
S_COEF = (-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
          2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10)
C_COEF = (4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
          -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11)


def det_sincos(x):
    """(sin x, cos x) for |x| up to a few thousand, about 1 ulp, as a fixed sequence of IEEE operations."""
    x = np.asarray(x, dtype=np.float64)
    kf = np.rint(x * TWO_OVER_PI)
    r = x - kf * PIO2_1
    r = r - kf * PIO2_2
    r = r - kf * PIO2_3
    r = r - kf * PIO2_3T
    z = r * r
    # sin r = r + r z (S1 + z (S2 + ... ))
    ps = S_COEF[5]
    for cf in S_COEF[4::-1]:
        ps = ps * z + cf
    sr = r + (r * z) * ps
    # cos r = 1 - z/2 + z z (C1 + z (C2 + ...))
    pc = C_COEF[5]
    for cf in C_COEF[4::-1]:
        pc = pc * z + cf
    cr = (1.0 - 0.5 * z) + (z * z) * pc
    q = kf.astype(np.int64) & 3
    s = np.where(q == 0, sr, np.where(q == 1, cr, np.where(q == 2, -sr, -cr)))
    c = np.where(q == 0, cr, np.where(q == 1, -sr, np.where(q == 2, -cr, sr)))
    return s, c


# ---- Clohessy-Wiltshire closed forms ---------------------------------------------------------------------------------
def cw_stm(T: float, nmm: float = 1.0) -> np.ndarray:
    """Phi(T) [6,6]; same entries as problems.cw_stm, each written as one fixed expression."""
    n = np.float64(nmm)
    T = np.float64(T)
    nT = n * T
    s, c = det_sincos(nT)
    s, c = np.float64(s), np.float64(c)
    omc = 1.0 - c
    P = np.zeros((6, 6))
    P[0, 0] = 4.0 - 3.0 * c
    P[1, 0] = 6.0 * (s - nT)
    P[1, 1] = 1.0
    P[2, 2] = c
    P[0, 3] = s / n
    P[0, 4] = (2.0 * omc) / n
    P[1, 3] = -((2.0 * omc) / n)
    P[1, 4] = (4.0 * s - 3.0 * nT) / n
    P[2, 5] = s / n
    P[3, 0] = (3.0 * n) * s
    P[4, 0] = -((6.0 * n) * omc)
    P[5, 2] = -(n * s)
    P[3, 3] = c
    P[3, 4] = 2.0 * s
    P[4, 3] = -(2.0 * s)
    P[4, 4] = 4.0 * c - 3.0
    P[5, 5] = c
    return P


def cw_zoh(T: float, nmm: float = 1.0):
    """(Phi(T), Gamma(T)) for a zero-order-hold acceleration; same entries as problems.cw_zoh."""
    n = np.float64(nmm)
    T = np.float64(T)
    nT = n * T
    s, c = det_sincos(nT)
    s, c = np.float64(s), np.float64(c)
    omc = 1.0 - c
    n2 = n * n
    Phi = cw_stm(T, nmm)
    G = np.zeros((6, 3))
    G[0, 0] = omc / n2
    G[0, 1] = (2.0 * (nT - s)) / n2
    G[1, 0] = -((2.0 * (nT - s)) / n2)
    G[1, 1] = (4.0 * omc - 1.5 * (nT * nT)) / n2
    G[2, 2] = omc / n2
    G[3, 0] = s / n
    G[3, 1] = (2.0 * omc) / n
    G[4, 0] = -((2.0 * omc) / n)
    G[4, 1] = (4.0 * s - 3.0 * nT) / n
    G[5, 2] = s / n
    return Phi, G


def cw_stage_matrices(kind: str, N: int, T: float, nmm: float = 1.0):
    """The shared (A, B) of the CW workloads, math layout (1,N,6,6) / (1,N,6,3).
    kind = 'cw_impulsive': A = Phi(T), B = Phi(T)[:, 3:6];  'cw_zoh': A = Phi(T), B = Gamma(T)."""
    if kind == "cw_impulsive":
        Phi = cw_stm(T, nmm)
        Gam = Phi[:, 3:6]
    elif kind == "cw_zoh":
        Phi, Gam = cw_zoh(T, nmm)
    else:
        raise ValueError(kind)
    return (np.broadcast_to(Phi, (1, N, 6, 6)).copy(), np.broadcast_to(Gam, (1, N, 6, 3)).copy())


# ---- elliptic orbit: RK4 of the LVLH linearised dynamics, zero-order-hold input --------------------------------------
def _elliptic_coeffs(th, e, p, h):
    """The seven non-trivial entries of Ac(theta) and theta' (vectorised over the batch)."""
    st, ct = det_sincos(th)
    one_ec = 1.0 + e * ct
    r = p / one_ec
    w = h / (r * r)
    rdot = (e * st) / h
    wdot = ((-2.0 * w) * rdot) / r
    k = 1.0 / ((r * r) * r)
    ww = w * w
    a30 = ww + 2.0 * k
    a41 = ww - k
    tw = 2.0 * w
    return a30, wdot, tw, a41, k, w      # Ac[3,0], Ac[3,1] = -Ac[4,0], Ac[3,4] = -Ac[4,3], Ac[4,1], -Ac[5,2], theta'


def _f_col(coef, y, forced_row):
    """dy = Ac y (+ 1 on `forced_row` for a Gamma column, else forced_row < 0); y: list of six (B,) arrays."""
    a30, wdot, tw, a41, k, _ = coef
    d3 = (a30 * y[0] + wdot * y[1]) + tw * y[4]
    d4 = ((-wdot) * y[0] + a41 * y[1]) + (-tw) * y[3]
    d5 = (-k) * y[2]
    dy = [y[3], y[4], y[5], d3, d4, d5]
    if forced_row >= 0:
        dy[forced_row] = dy[forced_row] + 1.0
    return dy


def elliptic_stage_matrices(e, theta0, N: int, T: float, substeps: int = 8):
    """Per-problem (Phi_k, Gamma_k), k = 0..N-1: classical RK4 (`substeps` per stage of length T) of
        y' = Ac(theta) y (+ e_{3+j} for Gamma column j),   theta' = h (1 + e cos theta)^2 / p^2 = w(theta)
    for each of the nine columns on its own (the device runs one thread per problem and column).
    -> A (B,N,6,6), B (B,N,6,3), math layout."""
    e = np.asarray(e, dtype=np.float64)
    th = np.array(theta0, dtype=np.float64)
    Bsz = e.shape[0]
    p = 1.0 - e * e
    h = np.sqrt(p)
    dt = np.float64(T) / np.float64(substeps)
    hdt = 0.5 * dt
    dt6 = dt / 6.0
    A = np.zeros((Bsz, N, 6, 6))
    Bm = np.zeros((Bsz, N, 6, 3))
    for kst in range(N):
        cols = []
        for j in range(9):
            y = [np.zeros(Bsz) for _ in range(6)]
            if j < 6:
                y[j] = np.ones(Bsz)
            cols.append(y)
        for _ in range(substeps):
            c1 = _elliptic_coeffs(th, e, p, h)
            th2 = th + hdt * c1[5]
            c2 = _elliptic_coeffs(th2, e, p, h)
            th3 = th + hdt * c2[5]
            c3 = _elliptic_coeffs(th3, e, p, h)
            th4 = th + dt * c3[5]
            c4 = _elliptic_coeffs(th4, e, p, h)
            for j in range(9):
                fr = j - 3 if j >= 6 else -1          # Gamma column j-6 is forced on row 3 + (j-6)
                y = cols[j]
                k1 = _f_col(c1, y, fr)
                k2 = _f_col(c2, [y[i] + hdt * k1[i] for i in range(6)], fr)
                k3 = _f_col(c3, [y[i] + hdt * k2[i] for i in range(6)], fr)
                k4 = _f_col(c4, [y[i] + dt * k3[i] for i in range(6)], fr)
                cols[j] = [y[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]) for i in range(6)]
            th = th + dt6 * (((c1[5] + 2.0 * c2[5]) + 2.0 * c3[5]) + c4[5])
        for j in range(6):
            for i in range(6):
                A[:, kst, i, j] = cols[j][i]
        for j in range(3):
            for i in range(6):
                Bm[:, kst, i, j] = cols[6 + j][i]
    return A, Bm
