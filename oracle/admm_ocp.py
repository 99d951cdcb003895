"""ORACLE (test infrastructure, NOT product code) -- NumPy restatement of oracle/admm_ocp.m.

PARITY STATUS: **unpinned by the reference**.  /root/reference/ contains only README.md:1-2
("Implementation of Alternating Direction Method of Multipliers for astrodynamics problems")
and LICENSE:1-21 -- no code, tests, fixtures or golden vectors.  BASELINE.json `north_star`
mandates for exactly this case "a minimal MATLAB ADMM written to the README's stated
algorithm ... committed as the correctness oracle and CPU baseline".  oracle/admm_ocp.m is
that artefact (MATLAB is absent from this image, so it is unexecuted here); this file is its
function-for-function NumPy restatement and the executable oracle; oracle/admm_ocp_cpu.c is
the C restatement with the canonical operation order (bit-comparable with the CUDA path).
The oracle itself is pinned only by independent cross-checks in tests/ (KKT residuals,
dense-KKT vs Riccati agreement, prox identities, HiGHS LP on config 1).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

Algorithm: scaled-form ADMM with over-relaxation (Boyd et al. 2011, sec. 3.1, 3.3, 3.4.1) on
    minimise 1/2 x'Px + q'x + g(z_J)   s.t.  Gx = h,  x_J = z_J
with J the index set of the SPLIT blocks (every block whose type is not BLK_NONE; BLK_FREE is a
split block with g = 0, BLK_NONE is a variable that takes no part in the splitting), x = (s_0,a_0,...,s_{N-1},a_{N-1},s_N), Gx = h the dynamics s_{k+1} = A_k s_k + B_k a_k + c_k,
s_0 = s_init, P = blkdiag(Q_0,R_0,...,Q_N), g block-separable over consecutive 3-vectors.
The x-update is written in the rho-scaled form
    x = argmin 1/2 x'(P/rho)x + (q/rho)'x + 1/2 |x_J - (z-u)_J|^2   s.t. Gx = h
so that for P = 0 the Riccati factor does not depend on rho (last
paragraph).  Rows of SURVEY.md section 8(a): a1 riccati_factor, a1' kkt_dense_factor,
a2 xupdate_riccati, a2' xupdate_dense, a3 prox_blocks, a4 dual_and_residuals, a5 adapt_rho,
a6 admm_solve.

All functions are batched: the leading axis is the problem index.  Shapes as in
admm-library_b200/problems.py.
"""
from __future__ import annotations

import numpy as np

BLK_FREE, BLK_L1, BLK_L1_BOX, BLK_L2, BLK_L2_BALL, BLK_BOX, BLK_BALL, BLK_POINT, BLK_NONE = range(9)
PAR_LAM, PAR_RAD, PAR_LO, PAR_HI = 0, 1, 2, 5

STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_NAN = 0, 1, 2
RHO_MAX, RHO_MIN = 1.0e6, 1.0e-6      # adaptive rho never leaves this range


# --------------------------------------------------------------------------- a1
def riccati_factor(A, B, c, Q, R, rho, wblk):
    """Row a1.  Backward Riccati recursion of the rho-scaled x-update.

    A (Bd,N,6,6)  B (Bd,N,6,3)  c (Bd,N,6)|None  Q (Bd,N+1,6,6)|None  R (Bd,N,3,3)|None
    rho (Bd,) or scalar; wblk (nb,) 0/1 split weight of every 3-block.  Returns per-stage arrays:
      K (Bd,N,3,6)  feedback gain           a_k = K_k s_k + d_k
      E (Bd,N,3,6)  Hinv B'                 d_k = Hinv ra_k + E_k g_k
      Hinv (Bd,N,3,3)
      Acl (Bd,N,6,6) A + B K                p_k = rs_k + K' ra_k + Acl' g_k
      chat (Bd,N,6)  P_{k+1} c_k            g_k = p_{k+1} - chat_k      (None if c is None)
      P (Bd,N+1,6,6) cost-to-go Hessians (scaled by 1/rho)
    """
    Bd, N = A.shape[0], A.shape[1]
    rho = np.broadcast_to(np.asarray(rho, dtype=np.float64), (Bd,))
    wblk = np.asarray(wblk, dtype=np.float64)

    def Ds(k):                                                    # split weights of state k
        return np.diag(np.repeat(wblk[3 * k:3 * k + 2], 3))[None]

    def Da(k):
        return (wblk[3 * k + 2] * np.eye(3))[None]

    K = np.zeros((Bd, N, 3, 6))
    E = np.zeros((Bd, N, 3, 6))
    Hinv = np.zeros((Bd, N, 3, 3))
    Acl = np.zeros((Bd, N, 6, 6))
    chat = None if c is None else np.zeros((Bd, N, 6))
    P = np.zeros((Bd, N + 1, 6, 6))
    Wn = np.broadcast_to(Ds(N), (Bd, 6, 6)) if Q is None else Ds(N) + Q[:, N] / rho[:, None, None]
    P[:, N] = Wn
    for k in range(N - 1, -1, -1):
        Pn = P[:, k + 1]
        Ak, Bk = A[:, k], B[:, k]
        BtP = np.einsum("bji,bjl->bil", Bk, Pn)                  # B' P+   (3x6)
        Wa = Da(k) if R is None else Da(k) + R[:, k] / rho[:, None, None]
        H = Wa + BtP @ Bk
        Hi = np.linalg.inv(H)
        Hi = 0.5 * (Hi + np.swapaxes(Hi, 1, 2))
        Kk = -Hi @ (BtP @ Ak)
        Ek = Hi @ np.swapaxes(Bk, 1, 2)
        Ac = Ak + Bk @ Kk
        Ws = Ds(k) if Q is None else Ds(k) + Q[:, k] / rho[:, None, None]
        Pk = Ws + np.swapaxes(Ak, 1, 2) @ Pn @ Ac
        P[:, k] = 0.5 * (Pk + np.swapaxes(Pk, 1, 2))
        K[:, k], E[:, k], Hinv[:, k], Acl[:, k] = Kk, Ek, Hi, Ac
        if c is not None:
            chat[:, k] = np.einsum("bij,bj->bi", Pn, c[:, k])
    return dict(K=K, E=E, Hinv=Hinv, Acl=Acl, chat=chat, P=P)


# --------------------------------------------------------------------------- a2
def xupdate_riccati(fac, A, B, c, s0, rt):
    """Row a2.  x = argmin of the scaled x-update for right-hand side rt = w*(z-u) - q/rho.

    rt (Bsz,n), s0 (Bsz,6); factor/dynamics with leading dim 1 (shared) or Bsz.  -> x (Bsz,n)."""
    Bsz, n = rt.shape
    N = (n - 6) // 9
    K, E, Hinv, Acl, chat = fac["K"], fac["E"], fac["Hinv"], fac["Acl"], fac["chat"]
    d = np.zeros((Bsz, N, 3))
    g = rt[:, 9 * N:9 * N + 6].copy()                            # p_N
    for k in range(N - 1, -1, -1):
        if chat is not None:
            g = g - chat[:, k]
        rs = rt[:, 9 * k:9 * k + 6]
        ra = rt[:, 9 * k + 6:9 * k + 9]
        d[:, k] = np.einsum("bij,bj->bi", Hinv[:, k], ra) + np.einsum("bij,bj->bi", E[:, k], g)
        g = rs + np.einsum("bji,bj->bi", K[:, k], ra) + np.einsum("bji,bj->bi", Acl[:, k], g)
    x = np.zeros((Bsz, n))
    s = np.array(s0, dtype=np.float64)
    for k in range(N):
        a = d[:, k] + np.einsum("bij,bj->bi", K[:, k], s)
        x[:, 9 * k:9 * k + 6] = s
        x[:, 9 * k + 6:9 * k + 9] = a
        s = np.einsum("bij,bj->bi", A[:, k], s) + np.einsum("bij,bj->bi", B[:, k], a)
        if c is not None:
            s = s + c[:, k]
    x[:, 9 * N:9 * N + 6] = s
    return x


# --------------------------------------------------------------------------- a1'
def assemble_G(A, B, N):
    """Dynamics constraint matrix of ONE problem: G (6(N+1) x n), rows [s_0 = s_init;
    s_{k+1} - A_k s_k - B_k a_k = c_k]."""
    n = 9 * N + 6
    G = np.zeros((6 * (N + 1), n))
    G[0:6, 0:6] = np.eye(6)
    for k in range(N):
        r = 6 * (k + 1)
        G[r:r + 6, 9 * k:9 * k + 6] = -A[k]
        G[r:r + 6, 9 * k + 6:9 * k + 9] = -B[k]
        G[r:r + 6, 9 * (k + 1):9 * (k + 1) + 6] = np.eye(6)
    return G


def assemble_P(Q, R, N):
    n = 9 * N + 6
    P = np.zeros((n, n))
    for k in range(N + 1):
        if Q is not None:
            P[9 * k:9 * k + 6, 9 * k:9 * k + 6] = Q[k]
        if R is not None and k < N:
            P[9 * k + 6:9 * k + 9, 9 * k + 6:9 * k + 9] = R[k]
    return P


def kkt_dense_factor(A, B, c, Q, R, rho, wblk):
    """Row a1'.  Shared-dynamics dense factor: with KKT = [[D + P/rho, G'],[G, 0]], D = diag(w),
    M = KKT^-1[:n,:n], S = KKT^-1[:n, n:n+6] (multiplies s_init), mc = KKT^-1[:n, n+6:] c.
    A (N,6,6) etc. for ONE (shared) model.  x = M rt + S s_init + mc."""
    N = A.shape[0]
    n = 9 * N + 6
    m = 6 * (N + 1)
    G = assemble_G(A, B, N)
    W = np.diag(np.repeat(np.asarray(wblk, dtype=np.float64), 3)) + assemble_P(Q, R, N) / rho
    KKT = np.zeros((n + m, n + m))
    KKT[:n, :n] = W
    KKT[:n, n:] = G.T
    KKT[n:, :n] = G
    Kinv = np.linalg.inv(KKT)
    M = Kinv[:n, :n]
    S = Kinv[:n, n:n + 6]
    mc = np.zeros(n) if c is None else Kinv[:n, n + 6:] @ c.reshape(-1)
    return dict(M=M, S=S, mc=mc, KKT=KKT)


# --------------------------------------------------------------------------- a2'
def xupdate_dense(dfac, s0, rt):
    """Row a2'.  X = M RT + S s_init + mc, one GEMM over the stacked right-hand sides."""
    return rt @ dfac["M"].T + s0 @ dfac["S"].T + dfac["mc"][None, :]


# --------------------------------------------------------------------------- a3
def prox_blocks(v, block_type, block_par, rinv):
    """Row a3.  z = prox_{g/rho}(v) block by block.  v (Bsz,n), block_par (Bp,nb,8),
    rinv = 1/rho (Bsz,)."""
    Bsz, n = v.shape
    nb = n // 3
    vb = v.reshape(Bsz, nb, 3)
    z = np.empty_like(vb)
    rinv = np.broadcast_to(np.asarray(rinv, dtype=np.float64), (Bsz,))
    for b in range(nb):
        t = int(block_type[b])
        par = block_par[:, b, :]                                  # (Bp,8)
        w = vb[:, b, :]
        kap = (par[:, PAR_LAM] * rinv)[:, None]
        lo = par[:, PAR_LO:PAR_LO + 3]
        hi = par[:, PAR_HI:PAR_HI + 3]
        rad = par[:, PAR_RAD]
        if t == BLK_FREE or t == BLK_NONE:
            z[:, b] = w
        elif t in (BLK_L1, BLK_L1_BOX):
            y = np.where(w > kap, w - kap, np.where(w < -kap, w + kap, 0.0))
            if t == BLK_L1_BOX:
                y = np.where(y < lo, lo, np.where(y > hi, hi, y))
            z[:, b] = y
        elif t in (BLK_L2, BLK_L2_BALL):
            nrm = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2])
            mag = nrm - kap[:, 0]
            if t == BLK_L2_BALL:
                mag = np.where(mag > rad, rad, mag)
            pos = nrm > kap[:, 0]
            scale = np.where(pos, mag / np.where(pos, nrm, 1.0), 0.0)
            z[:, b] = scale[:, None] * w
        elif t == BLK_BOX:
            z[:, b] = np.where(w < lo, lo, np.where(w > hi, hi, w))
        elif t == BLK_BALL:
            dv = w - lo                                           # centre stored in the lo slots
            nrm = np.sqrt(dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1] + dv[:, 2] * dv[:, 2])
            out = nrm > rad
            scale = np.where(out, rad / np.where(out, nrm, 1.0), 1.0)
            z[:, b] = np.where(out[:, None], lo + scale[:, None] * dv, w)
        elif t == BLK_POINT:
            z[:, b] = np.broadcast_to(lo, w.shape)
        else:
            raise ValueError(f"unknown block type {t}")
    return z.reshape(Bsz, n)


def split_weights(block_type):
    """(wblk (nb,), w (n,)): 1.0 on split blocks, 0.0 on BLK_NONE blocks."""
    wblk = (np.asarray(block_type) != BLK_NONE).astype(np.float64)
    return wblk, np.repeat(wblk, 3)


# --------------------------------------------------------------------------- a4
def dual_and_residuals(x, xh, z, z_old, u, rho, abstol, reltol, w):
    """Row a4.  Scaled dual ascent u += xh - z and the five norms / stopping thresholds, all
    restricted to the split entries (w = 1).  xh is the relaxed iterate alpha x + (1-alpha) z_old;
    z is already the prox output."""
    nsplit = float(np.sum(w))
    u_new = w * ((xh + u) - z)
    r_norm = np.sqrt(np.sum((w * (x - z)) ** 2, axis=1))
    s_norm = rho * np.sqrt(np.sum((w * (z - z_old)) ** 2, axis=1))
    eps_pri = np.sqrt(nsplit) * abstol + reltol * np.maximum(np.sqrt(np.sum((w * x) ** 2, axis=1)),
                                                             np.sqrt(np.sum((w * z) ** 2, axis=1)))
    eps_dual = np.sqrt(nsplit) * abstol + reltol * rho * np.sqrt(np.sum(u_new * u_new, axis=1))
    return u_new, r_norm, s_norm, eps_pri, eps_dual


# --------------------------------------------------------------------------- a5
def adapt_rho(r_norm, s_norm, rho, mu, tau):
    """Row a5.  Residual balancing (Boyd 2011 eq. 3.13).  Returns (rho_new, u_scale) with
    u_scale the factor the scaled dual must be multiplied by (rho_old / rho_new)."""
    inv_tau = 1.0 / tau
    want_up = r_norm > mu * s_norm
    up = want_up & ~(rho * tau > RHO_MAX)
    dn = (~want_up) & (s_norm > mu * r_norm) & ~(rho * inv_tau < RHO_MIN)
    rho_new = np.where(up, rho * tau, np.where(dn, rho * inv_tau, rho))
    u_scale = np.where(up, inv_tau, np.where(dn, tau, 1.0))
    return rho_new, u_scale


# --------------------------------------------------------------------------- a6
def admm_solve(prob, opts):
    """Row a6.  [x, z, u, hist] = admm_solve(prob, opts): the reference-facing surface.

    Outputs on BLK_NONE entries: z = x, u = 0 (they are not part of the splitting)."""
    A, B, c, Q, R, q = prob["A"], prob["B"], prob.get("c"), prob.get("Q"), prob.get("R"), prob.get("q")
    s0 = np.asarray(prob["s0"], dtype=np.float64)
    bt, bp = prob["block_type"], prob["block_par"]
    Bsz = s0.shape[0]
    N = A.shape[1]
    n = 9 * N + 6
    Bd = A.shape[0]
    alpha = float(opts.get("alpha", 1.0))
    abstol = float(opts.get("abstol", 1e-6))
    reltol = float(opts.get("reltol", 1e-6))
    max_iter = int(opts.get("max_iter", 1000))
    adapt = bool(opts.get("adapt_rho", 0))
    mu, tau = float(opts.get("adapt_mu", 10.0)), float(opts.get("adapt_tau", 2.0))
    every = int(opts.get("adapt_every", 25))
    until = int(opts.get("adapt_until", 0))
    want_hist = bool(opts.get("history", 0))
    xupd = opts.get("xupdate", "auto")
    rho = np.full(Bsz, float(opts.get("rho", 1.0))) if prob.get("rho0") is None \
        else np.array(prob["rho0"], dtype=np.float64)
    wblk, w = split_weights(bt)
    z = np.zeros((Bsz, n)) if prob.get("z0") is None else np.array(prob["z0"], dtype=np.float64)
    u = np.zeros((Bsz, n)) if prob.get("u0") is None else np.array(prob["u0"], dtype=np.float64)
    z, u = w * z, w * u
    has_P = (Q is not None) or (R is not None)

    use_dense = (xupd == "dense")
    if use_dense:
        if Bd != 1:
            raise ValueError("dense x-update needs shared dynamics")
        if has_P and (adapt or prob.get("rho0") is not None):
            raise ValueError("dense x-update with P != 0 needs one shared rho")
        dfac = kkt_dense_factor(A[0], B[0], None if c is None else c[0],
                                None if Q is None else Q[0], None if R is None else R[0],
                                rho[0], wblk)
    else:
        # the factor depends on rho only through P/rho
        per_rho = adapt or prob.get("rho0") is not None
        if has_P and Bd == 1 and per_rho and Bsz > 1:
            # per-problem rho needs a per-problem factor: broadcast the model
            Af = np.broadcast_to(A, (Bsz,) + A.shape[1:])
            Bf = np.broadcast_to(B, (Bsz,) + B.shape[1:])
            cf = None if c is None else np.broadcast_to(c, (Bsz,) + c.shape[1:])
            Qf = None if Q is None else np.broadcast_to(Q, (Bsz,) + Q.shape[1:])
            Rf = None if R is None else np.broadcast_to(R, (Bsz,) + R.shape[1:])
        else:
            Af, Bf, cf, Qf, Rf = A, B, c, Q, R
        rho_f = rho if Af.shape[0] == Bsz else rho[:1]
        fac = riccati_factor(Af, Bf, cf, Qf, Rf, rho_f, wblk)

    iters = np.zeros(Bsz, dtype=np.int32)
    status = np.full(Bsz, STATUS_MAX_ITER, dtype=np.int32)
    active = np.ones(Bsz, dtype=bool)
    x = np.zeros((Bsz, n))
    fin = {k: np.zeros(Bsz) for k in ("r_norm", "s_norm", "eps_pri", "eps_dual")}
    hist = {k: np.full((Bsz, max_iter), np.nan) for k in
            ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho")} if want_hist else {}
    refactor_count = 0

    for k in range(1, max_iter + 1):
        if not active.any():
            break
        rinv = 1.0 / rho
        rt = w * (z - u)
        if q is not None:
            rt = rt - q * rinv[:, None]
        if use_dense:
            x_new = xupdate_dense(dfac, s0, rt)
        else:
            x_new = xupdate_riccati(fac, Af, Bf, cf, s0, rt)
        xh = alpha * x_new + (1.0 - alpha) * z
        z_new = w * prox_blocks(xh + u, bt, bp, rinv)
        u_new, r_norm, s_norm, eps_pri, eps_dual = dual_and_residuals(
            x_new, xh, z_new, z, u, rho, abstol, reltol, w)
        a = active
        x[a], z[a], u[a] = x_new[a], z_new[a], u_new[a]
        iters[a] = k
        for name, val in (("r_norm", r_norm), ("s_norm", s_norm),
                          ("eps_pri", eps_pri), ("eps_dual", eps_dual)):
            fin[name][a] = val[a]
            if want_hist:
                hist[name][a, k - 1] = val[a]
        if want_hist:
            hist["rho"][a, k - 1] = rho[a]
        bad = a & ~(np.isfinite(r_norm) & np.isfinite(s_norm))
        conv = a & ~bad & (r_norm < eps_pri) & (s_norm < eps_dual)
        status[conv] = STATUS_CONVERGED
        status[bad] = STATUS_NAN
        active = a & ~conv & ~bad
        if adapt and (k % every == 0) and k < max_iter and (until <= 0 or k <= until) and active.any():
            rho_new, usc = adapt_rho(r_norm, s_norm, rho, mu, tau)
            ch = active & (rho_new != rho)
            rho = np.where(ch, rho_new, rho)
            u = np.where(ch[:, None], u * usc[:, None], u)
            if ch.any() and has_P and not use_dense:
                refactor_count += int(ch.sum())
                fac = riccati_factor(Af, Bf, cf, Qf, Rf, rho, wblk)  # only rows in `ch` change
    z = z + (1.0 - w) * x                                         # BLK_NONE entries: z = x, u = 0
    out_hist = dict(iters=iters, status=status, rho=rho, refactor_count=refactor_count, **fin)
    if want_hist:
        out_hist["hist"] = hist
    return x, z, u, out_hist


# --------------------------------------------------------------------------- helpers for tests
def objective(prob, x):
    """Objective 1/2 x'Px + q'x + g(x) (indicator terms must be satisfied; not checked)."""
    Bsz, n = x.shape
    N = (n - 6) // 9
    val = np.zeros(Bsz)
    Q, R, q = prob.get("Q"), prob.get("R"), prob.get("q")
    for k in range(N + 1):
        if Q is not None:
            s = x[:, 9 * k:9 * k + 6]
            val += 0.5 * np.einsum("bi,bij,bj->b", s, np.broadcast_to(Q[:, k], (Bsz, 6, 6)), s)
        if R is not None and k < N:
            a = x[:, 9 * k + 6:9 * k + 9]
            val += 0.5 * np.einsum("bi,bij,bj->b", a, np.broadcast_to(R[:, k], (Bsz, 3, 3)), a)
    if q is not None:
        val += np.sum(q * x, axis=1)
    bt, bp = prob["block_type"], prob["block_par"]
    xb = x.reshape(Bsz, n // 3, 3)
    for b, t in enumerate(bt):
        lam = bp[:, b, PAR_LAM]
        if t in (BLK_L1, BLK_L1_BOX):
            val += lam * np.sum(np.abs(xb[:, b]), axis=1)
        elif t in (BLK_L2, BLK_L2_BALL):
            val += lam * np.sqrt(np.sum(xb[:, b] ** 2, axis=1))
    return val
