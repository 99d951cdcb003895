"""ORACLE (test infrastructure) -- ctypes wrapper around oracle/_build/libadmm_ocp_cpu.so, the C
restatement with the canonical operation order (see admm_ocp_cpu.c header; parity unpinned by the
reference, which contains no code).  Takes the same Python-side problem dict as admm_ocp.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libadmm_ocp_cpu.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OcpProblem(C.Structure):
    _fields_ = [("N", C.c_int32), ("batch", C.c_int64),
                ("A", c_dp), ("dyn_batched", C.c_int32), ("B", c_dp), ("c", c_dp), ("Q", c_dp),
                ("R", c_dp), ("q", c_dp), ("q_batched", C.c_int32), ("s0", c_dp),
                ("block_type", c_ip), ("block_par", c_dp), ("par_batched", C.c_int32),
                ("z0", c_dp), ("u0", c_dp), ("rho0", c_dp)]


class OcpOpts(C.Structure):
    _fields_ = [("rho", C.c_double), ("alpha", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("max_iter", C.c_int32), ("adapt_rho", C.c_int32),
                ("adapt_mu", C.c_double), ("adapt_tau", C.c_double), ("adapt_every", C.c_int32), ("adapt_until", C.c_int32),
                ("xupdate", C.c_int32), ("history", C.c_int32)]


class OcpResult(C.Structure):
    _fields_ = [("x", c_dp), ("z", c_dp), ("u", c_dp), ("iters", c_ip), ("status", c_ip),
                ("r_norm", c_dp), ("s_norm", c_dp), ("eps_pri", c_dp), ("eps_dual", c_dp),
                ("rho", c_dp),
                ("hist_r", c_dp), ("hist_s", c_dp), ("hist_eps_pri", c_dp), ("hist_eps_dual", c_dp),
                ("hist_rho", c_dp), ("stats", C.c_int64 * 4), ("seconds", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "admm_ocp_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ocp_solve.restype = C.c_int
        _lib.ocp_solve.argtypes = [C.POINTER(OcpProblem), C.POINTER(OcpOpts), C.POINTER(OcpResult), C.c_int]
        _lib.ocp_factor_stride.restype = C.c_int
        _lib.ocp_num_threads.restype = C.c_int
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def to_matlab_layout(prob: dict) -> dict:
    """Python math-layout arrays -> contiguous MATLAB column-major buffers (the C-ABI layout)."""
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    t = lambda a: None if a is None else np.ascontiguousarray(  # noqa: E731
        np.swapaxes(np.asarray(a, dtype=np.float64), -1, -2))
    Bsz = prob["s0"].shape[0]
    out = dict(N=int(prob["A"].shape[1]), batch=Bsz,
               A=t(prob["A"]), B=t(prob["B"]), c=f(prob.get("c")), Q=t(prob.get("Q")),
               R=t(prob.get("R")), q=f(prob.get("q")), s0=f(prob["s0"]),
               block_type=np.ascontiguousarray(prob["block_type"], dtype=np.int32),
               block_par=f(prob["block_par"]), z0=f(prob.get("z0")), u0=f(prob.get("u0")),
               rho0=f(prob.get("rho0")))
    out["dyn_batched"] = int(prob["A"].shape[0] > 1 or (prob["A"].shape[0] == Bsz and Bsz == 1 and False))
    if prob["A"].shape[0] not in (1, Bsz):
        raise ValueError("A leading dim must be 1 or batch")
    for k in ("B", "c", "Q", "R"):
        a = prob.get(k)
        if a is not None and a.shape[0] != prob["A"].shape[0]:
            raise ValueError(f"{k} must be batched like A")
    out["q_batched"] = int(prob.get("q") is not None and prob["q"].shape[0] > 1)
    out["par_batched"] = int(prob["block_par"].shape[0] > 1)
    return out


XUPDATE = {"auto": 0, "dense": 1, "riccati": 2}


def solve(prob: dict, opts: dict, nthreads: int = 0):
    """[x, z, u, hist] = solve(prob, opts) with the C oracle.  Same outputs as admm_ocp.admm_solve."""
    L = lib()
    m = to_matlab_layout(prob)
    N, Bsz = m["N"], m["batch"]
    n = 9 * N + 6
    pb = OcpProblem(N=N, batch=Bsz, A=_dp(m["A"]), dyn_batched=m["dyn_batched"], B=_dp(m["B"]),
                    c=_dp(m["c"]), Q=_dp(m["Q"]), R=_dp(m["R"]), q=_dp(m["q"]),
                    q_batched=m["q_batched"], s0=_dp(m["s0"]),
                    block_type=m["block_type"].ctypes.data_as(c_ip), block_par=_dp(m["block_par"]),
                    par_batched=m["par_batched"], z0=_dp(m["z0"]), u0=_dp(m["u0"]), rho0=_dp(m["rho0"]))
    max_iter = int(opts.get("max_iter", 1000))
    op = OcpOpts(rho=float(opts.get("rho", 1.0)), alpha=float(opts.get("alpha", 1.0)),
                 abstol=float(opts.get("abstol", 1e-6)), reltol=float(opts.get("reltol", 1e-6)),
                 max_iter=max_iter, adapt_rho=int(opts.get("adapt_rho", 0)),
                 adapt_mu=float(opts.get("adapt_mu", 10.0)), adapt_tau=float(opts.get("adapt_tau", 2.0)),
                 adapt_every=int(opts.get("adapt_every", 25)), adapt_until=int(opts.get("adapt_until", 0)),
                 xupdate=XUPDATE[opts.get("xupdate", "auto")], history=int(opts.get("history", 0)))
    x = np.empty((Bsz, n)); z = np.empty((Bsz, n)); u = np.empty((Bsz, n))
    iters = np.zeros(Bsz, dtype=np.int32); status = np.zeros(Bsz, dtype=np.int32)
    fin = {k: np.zeros(Bsz) for k in ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho")}
    res = OcpResult(x=_dp(x), z=_dp(z), u=_dp(u), iters=iters.ctypes.data_as(c_ip),
                    status=status.ctypes.data_as(c_ip), **{k: _dp(v) for k, v in fin.items()})
    hist = None
    if op.history:
        hist = {k: np.empty((Bsz, max_iter)) for k in ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho")}
        res.hist_r, res.hist_s = _dp(hist["r_norm"]), _dp(hist["s_norm"])
        res.hist_eps_pri, res.hist_eps_dual = _dp(hist["eps_pri"]), _dp(hist["eps_dual"])
        res.hist_rho = _dp(hist["rho"])
    rc = L.ocp_solve(C.byref(pb), C.byref(op), C.byref(res), int(nthreads))
    if rc < 0:
        raise ValueError("oracle rejected the problem/option combination")
    out = dict(iters=iters, status=status, stats=list(res.stats), seconds=res.seconds,
               refactor_count=int(res.stats[3]), **fin)
    if hist is not None:
        out["hist"] = hist
    return x, z, u, out
