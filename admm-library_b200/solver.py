"""Host-side mirror of the MATLAB surface  [x, z, u, hist] = admm_solve(prob, opts)
(admm-library_b200/matlab/admm_solve.m) over the C ABI.

Python arrays use the "math" layout of problems.py; this module converts them to the MATLAB
column-major buffers the C ABI takes, calls libadmm_b200.so and returns NumPy arrays.  No numeric
work happens here and there is no fallback: every call goes to the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

XUPDATE = {"auto": 0, "dense": 1, "riccati": 2}
# tf32: tensor cores allowed (3xTF32 split GEMM on the x-update increments; with xupdate="dense" throughout, with
# "auto" once the working set is narrow); tf32_single: unit entry point only
PRECISION = {"fp64": 0, "tf32": 1, "tf32_single": 2}
# kernel variant of the FP64 Riccati path (include/admm_b200.h ADMMB_KERNEL_*): "auto" in normal use
KERNEL = {"auto": 0, "thread": 1, "thread_wide": 2, "thread2": 3, "tile": 4, "wg": 5, "pint": 6}
DEFAULT_KERNEL = "auto"   # what make_opts uses when opts has no "kernel" key (the GPU test suite pins each variant in turn)
FS = 156  # doubles per stage in a factor record (csrc/common.cuh)
# name -> (offset, rows, cols, row stride) inside one record
FAC_LAYOUT = dict(K=(0, 3, 6, 6), Acl=(18, 6, 6, 6), Hinv=(54, 3, 3, 4), E=(66, 3, 6, 6), A=(84, 6, 6, 6),
                  B=(120, 6, 3, 4), c=(144, 1, 6, 6), chat=(150, 1, 6, 6))


def unpack_factor(fac, layout=FAC_LAYOUT):
    """factor records [N, FS] -> dict of [N, rows, cols] arrays."""
    out = {}
    for name, (off, r, c, ldr) in layout.items():
        out[name] = np.stack([fac[:, off + i * ldr: off + i * ldr + c] for i in range(r)], axis=1)
    return out


def pack_factor(parts, N, layout=FAC_LAYOUT, fs=FS):
    fac = np.zeros((N, fs))
    for name, (off, r, c, ldr) in layout.items():
        for i in range(r):
            fac[:, off + i * ldr: off + i * ldr + c] = parts[name][:, i, :]
    return fac


GENERATOR = {"cw_impulsive": 1, "cw_zoh": 2, "elliptic_zoh": 3}   # include/admm_b200.h ADMMB_GEN_*


def make_generator(gen: dict):
    """gen = dict(kind=..., T=..., nmm=1.0, substeps=8, e=[batch], theta0=[batch]) -> (ctypes struct, buffers to keep alive)."""
    e = None if gen.get("e") is None else np.ascontiguousarray(gen["e"], dtype=np.float64)
    th = None if gen.get("theta0") is None else np.ascontiguousarray(gen["theta0"], dtype=np.float64)
    g = L.Generator(kind=GENERATOR[gen["kind"]], substeps=int(gen.get("substeps", 0)), T=float(gen["T"]),
                    nmm=float(gen.get("nmm", 0.0)), e=_dp(e), theta0=_dp(th))
    return g, (e, th)


SCP_MODEL = {"nl_circular": 1, "nl_elliptic": 2}   # include/admm_b200.h ADMMB_SCP_*
SCP_CONTROL = {"zoh": 0, "impulsive": 1}   # ADMMB_SCP_CTRL_*


def make_scp(scp: dict) -> L.Scp:
    """scp = dict(T=..., R0=..., max_pass=..., tol_abs=..., tol_rel=..., model="nl_circular", nmm=1.0, substeps=8,
    control="zoh" | "impulsive"; model="nl_elliptic" adds e=[batch], theta0=[batch])."""
    e = None if scp.get("e") is None else np.ascontiguousarray(scp["e"], dtype=np.float64)
    th = None if scp.get("theta0") is None else np.ascontiguousarray(scp["theta0"], dtype=np.float64)
    sc = L.Scp(model=SCP_MODEL[scp.get("model", "nl_circular")], substeps=int(scp.get("substeps", 0)),
               T=float(scp["T"]), nmm=float(scp.get("nmm", 0.0)), R0=float(scp["R0"]), max_pass=int(scp["max_pass"]),
               tol_abs=float(scp.get("tol_abs", 0.0)), tol_rel=float(scp.get("tol_rel", 0.0)),
               control=SCP_CONTROL[scp.get("control", "zoh")], e=_dp(e), theta0=_dp(th))
    sc._keep = (e, th)        # the struct holds raw pointers into these
    return sc


def _dp(a):
    return None if a is None else a.ctypes.data_as(L.c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(L.c_ip)


def to_c_layout(prob: dict) -> dict:
    """math-layout dict -> contiguous MATLAB column-major buffers + flags."""
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    t = lambda a: None if a is None else np.ascontiguousarray(  # noqa: E731
        np.swapaxes(np.asarray(a, dtype=np.float64), -1, -2))
    Bsz = int(prob["s0"].shape[0])
    if prob.get("A") is None:     # generated model (Solver.solve_generated): A, B are computed on the device
        Bd, N = 1, int(prob["N"])
    else:
        Bd, N = int(prob["A"].shape[0]), int(prob["A"].shape[1])
    if Bd not in (1, Bsz):
        raise ValueError("A leading dimension must be 1 (shared) or batch")
    for k in ("B", "c", "Q", "R"):
        a = prob.get(k)
        if a is not None and prob.get("A") is not None and a.shape[0] != Bd:
            raise ValueError(f"{k} must be batched like A")
    m = dict(N=N, batch=Bsz, dyn_batched=int(Bd > 1),
             A=t(prob.get("A")), B=t(prob.get("B")), c=f(prob.get("c")), Q=t(prob.get("Q")), R=t(prob.get("R")),
             q=f(prob.get("q")), s0=f(prob["s0"]),
             block_type=np.ascontiguousarray(prob["block_type"], dtype=np.int32),
             block_par=f(prob["block_par"]), z0=f(prob.get("z0")), u0=f(prob.get("u0")),
             rho0=f(prob.get("rho0")))
    m["q_batched"] = int(m["q"] is not None and m["q"].shape[0] > 1)
    m["par_batched"] = int(m["block_par"].shape[0] > 1)
    return m


def make_problem(m: dict) -> L.Problem:
    return L.Problem(N=m["N"], batch=m["batch"], A=_dp(m["A"]), dyn_batched=m["dyn_batched"], B=_dp(m["B"]),
                     c=_dp(m["c"]), Q=_dp(m["Q"]), R=_dp(m["R"]), q=_dp(m["q"]), q_batched=m["q_batched"],
                     s0=_dp(m["s0"]), block_type=_ip(m["block_type"]), block_par=_dp(m["block_par"]),
                     par_batched=m["par_batched"], z0=_dp(m["z0"]), u0=_dp(m["u0"]), rho0=_dp(m["rho0"]))


def make_opts(opts: dict) -> L.Opts:
    return L.Opts(rho=float(opts.get("rho", 1.0)), alpha=float(opts.get("alpha", 1.0)),
                  abstol=float(opts.get("abstol", 1e-6)), reltol=float(opts.get("reltol", 1e-6)),
                  max_iter=int(opts.get("max_iter", 1000)), adapt_rho=int(opts.get("adapt_rho", 0)),
                  adapt_mu=float(opts.get("adapt_mu", 10.0)), adapt_tau=float(opts.get("adapt_tau", 2.0)),
                  adapt_every=int(opts.get("adapt_every", 25)), adapt_until=int(opts.get("adapt_until", 0)),
                  xupdate=XUPDATE[opts.get("xupdate", "auto")], precision=PRECISION[opts.get("precision", "fp64")],
                  history=int(opts.get("history", 0)), chunk=int(opts.get("chunk", 0)),
                  kernel=KERNEL[opts.get("kernel", DEFAULT_KERNEL)], tf32_switch=int(opts.get("tf32_switch", 0)),
                  tf32_refresh=int(opts.get("tf32_refresh", 0)))


class ResultBuffers:
    """Caller-owned output buffers (NumPy or any object exposing .ctypes / data pointers)."""

    def __init__(self, batch: int, n: int, max_iter: int, history: bool, want=("x", "z", "u"), alloc=None):
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))
        self.x = alloc((batch, n), np.float64) if "x" in want else None
        self.z = alloc((batch, n), np.float64) if "z" in want else None
        self.u = alloc((batch, n), np.float64) if "u" in want else None
        self.iters = np.zeros(batch, dtype=np.int32)
        self.status = np.zeros(batch, dtype=np.int32)
        self.fin = {k: np.zeros(batch) for k in ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho")}
        self.hist = None
        if history:
            self.hist = {k: np.full((batch, max_iter), np.nan) for k in
                         ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho")}
        self.c = L.Result(x=_dp(self.x), z=_dp(self.z), u=_dp(self.u), iters=_ip(self.iters),
                          status=_ip(self.status), **{k: _dp(v) for k, v in self.fin.items()})
        if history:
            self.c.hist_r, self.c.hist_s = _dp(self.hist["r_norm"]), _dp(self.hist["s_norm"])
            self.c.hist_eps_pri, self.c.hist_eps_dual = _dp(self.hist["eps_pri"]), _dp(self.hist["eps_dual"])
            self.c.hist_rho = _dp(self.hist["rho"])

    def hist_dict(self) -> dict:
        out = dict(iters=self.iters, status=self.status, stats=list(self.c.stats),
                   refactor_count=int(self.c.stats[3]), device_ms=self.c.device_ms, h2d_ms=self.c.h2d_ms,
                   d2h_ms=self.c.d2h_ms, launches=int(self.c.launches), kernel_ms=self.c.kernel_ms,
                   kernel_launches=int(self.c.kernel_launches), **self.fin)
        if self.hist is not None:
            # iterations a problem never ran stay NaN, as in the oracle
            k = np.arange(self.hist["r_norm"].shape[1])[None, :]
            mask = k >= self.iters[:, None]
            for v in self.hist.values():
                v[mask] = np.nan
            out["hist"] = self.hist
        return out


class Solver:
    """Owns an admmb_handle (device memory, streams, one worker thread per GPU)."""

    def __init__(self, devices=None):
        self._L = L.load()
        self._h = C.c_void_p()
        if devices is None:
            ids, nd = None, 1
        elif isinstance(devices, int):
            ids, nd = None, devices
        else:
            ids, nd = (C.c_int * len(devices))(*devices), len(devices)
        rc = self._L.admmb_create(C.byref(self._h), ids, nd)
        if rc != L.OK:
            msg = self._L.admmb_last_error(None)
            raise L.AdmmError(rc, msg.decode() if msg else "")
        self._keep = None
        self._up = None
        self._n = self._batch = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.admmb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != L.OK:
            msg = self._L.admmb_last_error(self._h)
            raise L.AdmmError(rc, msg.decode() if msg else "")

    @property
    def device_count(self) -> int:
        return int(self._L.admmb_device_count(self._h))

    @property
    def nccl_gathers(self) -> int:
        """Statistics all-reduces done over NCCL by this (multi-GPU) handle; 0: one GPU or libnccl not found."""
        return int(self._L.admmb_nccl_gathers(self._h))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._L.admmb_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    # ---- one-shot -------------------------------------------------------------------------
    def solve(self, prob: dict, opts: dict, want=("x", "z", "u")):
        """[x, z, u, hist] = solve(prob, opts) -- host buffers in, host buffers out."""
        m = to_c_layout(prob)
        n = 9 * m["N"] + 6
        op = make_opts(opts)
        res = ResultBuffers(m["batch"], n, op.max_iter, bool(op.history), want)
        pb = make_problem(m)
        self._check(self._L.admmb_solve(self._h, C.byref(pb), C.byref(op), C.byref(res.c)))
        return res.x, res.z, res.u, res.hist_dict()

    def solve_generated(self, prob: dict, gen: dict, opts: dict, want=("x", "z", "u")):
        """solve() with the stage matrices generated on the device: prob carries N instead of A, B."""
        m = to_c_layout(prob)
        n = 9 * m["N"] + 6
        op = make_opts(opts)
        res = ResultBuffers(m["batch"], n, op.max_iter, bool(op.history), want)
        pb = make_problem(m)
        g, keep = make_generator(gen)
        self._check(self._L.admmb_solve_generated(self._h, C.byref(pb), C.byref(g), C.byref(op), C.byref(res.c)))
        return res.x, res.z, res.u, res.hist_dict()

    def scp_solve(self, prob: dict, scp: dict, opts: dict, want=("x", "z", "u")):
        """Sequential convex programming on the device (oracle/scp_ocp.py scp_solve): prob carries N, s0,
        the block table and optionally q / per-problem Q, R -- the stage records are re-linearised on the GPU each pass.
        -> x, z, u, info (hist_dict of the last convex solves + passes, scp_status, step, iters_total, hist_step)."""
        m = to_c_layout(dict(prob, A=None, B=None, c=None))
        n = 9 * m["N"] + 6
        Bsz = m["batch"]
        op = make_opts(opts)
        sc = make_scp(scp)
        res = ResultBuffers(Bsz, n, op.max_iter, False, want)
        pb = make_problem(m)
        passes = np.zeros(Bsz, dtype=np.int32)
        scp_status = np.zeros(Bsz, dtype=np.int32)
        step = np.zeros(Bsz)
        iters_total = np.zeros(Bsz, dtype=np.int64)
        hist_step = np.zeros((Bsz, sc.max_pass))
        out = L.ScpResult(passes=_ip(passes), scp_status=_ip(scp_status), step=_dp(step),
                          iters_total=iters_total.ctypes.data_as(C.POINTER(C.c_int64)), hist_step=_dp(hist_step))
        self._check(self._L.admmb_scp_solve(self._h, C.byref(pb), C.byref(sc), C.byref(op), C.byref(res.c), C.byref(out)))
        info = res.hist_dict()
        info.update(passes=passes, scp_status=scp_status, step=step, iters_total=iters_total, hist_step=hist_step,
                    scp_stats=list(out.stats), linearise_ms=out.linearise_ms)
        return res.x, res.z, res.u, info

    def k_scp_linearise(self, N: int, scp: dict, xref: np.ndarray, s0: np.ndarray | None = None):
        """One linearisation pass in math layout: about xref [B, n] (s0 None) or along the nonlinear trajectory from s0
        under xref's controls (-> that trajectory as well).  -> A (B,N,6,6), B (B,N,6,3), c (B,N,6), xref."""
        xr = np.array(xref, dtype=np.float64, order="C")
        Bsz = xr.shape[0]
        A = np.zeros((Bsz, N, 6, 6))
        Bm = np.zeros((Bsz, N, 3, 6))
        c = np.zeros((Bsz, N, 6))
        s0c = None if s0 is None else np.ascontiguousarray(s0, dtype=np.float64)
        sc = make_scp(dict(scp, max_pass=scp.get("max_pass", 1)))
        self._check(self._L.admmb_k_scp_linearise(self._h, int(N), Bsz, C.byref(sc), int(s0 is not None), _dp(s0c),
                                                  _dp(xr), _dp(A), _dp(Bm), _dp(c)))
        return (np.ascontiguousarray(np.swapaxes(A, -1, -2)), np.ascontiguousarray(np.swapaxes(Bm, -1, -2)), c, xr)

    def upload_generated(self, prob: dict, gen: dict, opts: dict):
        m = to_c_layout(prob)
        self._keep = m
        self._n = 9 * m["N"] + 6
        self._batch = m["batch"]
        pb = make_problem(m)
        op = make_opts(opts)
        g, keep = make_generator(gen)
        self._check(self._L.admmb_upload_generated(self._h, C.byref(pb), C.byref(g), C.byref(op)))
        self._up = (int(op.max_iter), bool(op.history))

    def k_generate(self, N: int, batch: int, gen: dict):
        """The generated stage matrices in math layout: A (Bd,N,6,6), B (Bd,N,6,3)."""
        g, keep = make_generator(gen)
        Bd = batch if gen["kind"] == "elliptic_zoh" else 1
        A = np.zeros((Bd, N, 6, 6))     # column-major 6x6 per stage on the C side: transposed below
        B = np.zeros((Bd, N, 3, 6))
        self._check(self._L.admmb_k_generate(self._h, int(N), int(batch), C.byref(g), _dp(A), _dp(B)))
        return np.ascontiguousarray(np.swapaxes(A, -1, -2)), np.ascontiguousarray(np.swapaxes(B, -1, -2))

    # ---- staged ----------------------------------------------------------------------------
    def upload(self, prob: dict, opts: dict):
        m = to_c_layout(prob)
        self._keep = m
        self._n = 9 * m["N"] + 6
        self._batch = m["batch"]
        pb = make_problem(m)
        op = make_opts(opts)
        self._check(self._L.admmb_upload(self._h, C.byref(pb), C.byref(op)))
        self._up = (int(op.max_iter), bool(op.history))     # what the library sized its history buffers with

    def upload_c(self, pb: L.Problem, op: L.Opts, batch: int, n: int):
        """Upload from caller-built ctypes structs (e.g. pointing into pinned memory)."""
        self._n, self._batch = n, batch
        self._check(self._L.admmb_upload(self._h, C.byref(pb), C.byref(op)))
        self._up = (int(op.max_iter), bool(op.history))

    def run(self, opts: dict | L.Opts) -> dict:
        op = opts if isinstance(opts, L.Opts) else make_opts(opts)
        r = L.Result()
        self._check(self._L.admmb_run(self._h, C.byref(op), C.byref(r)))
        return dict(stats=list(r.stats), device_ms=r.device_ms, launches=int(r.launches),
                    kernel_ms=r.kernel_ms, kernel_launches=int(r.kernel_launches))

    def shift_resolve(self, k: int, opts: dict | L.Opts, s0_new=None) -> dict:
        """Receding-horizon step on the resident batch: warm start = the last solution shifted by k stages, initial states
        = s0_new [batch, 6] (None: the solution's state at stage k); solves again without re-uploading the model."""
        op = opts if isinstance(opts, L.Opts) else make_opts(opts)
        r = L.Result()
        s0c = None if s0_new is None else np.ascontiguousarray(s0_new, dtype=np.float64)
        self._check(self._L.admmb_shift_resolve(self._h, int(k), _dp(s0c), C.byref(op), C.byref(r)))
        return dict(stats=list(r.stats), device_ms=r.device_ms, launches=int(r.launches),
                    kernel_ms=r.kernel_ms, kernel_launches=int(r.kernel_launches))

    def download(self, opts: dict | None = None, want=("x", "z", "u")):
        # the history buffers are sized by the UPLOADED options (the library writes max_iter_uploaded x batch values)
        max_iter, history = getattr(self, "_up", None) or (make_opts(opts).max_iter, bool(make_opts(opts).history))
        res = ResultBuffers(max(self._batch, 1), max(self._n, 1), max_iter, history, want)
        self._check(self._L.admmb_download(self._h, C.byref(res.c)))
        return res.x, res.z, res.u, res.hist_dict()

    def download_c(self, res: L.Result):
        self._check(self._L.admmb_download(self._h, C.byref(res)))

    # ---- unit kernels (tests) -----------------------------------------------------------------
    def k_riccati_factor(self, prob: dict, rho: float) -> np.ndarray:
        m = to_c_layout(prob)
        assert m["dyn_batched"] == 0
        fac = np.zeros((m["N"], FS))
        self._check(self._L.admmb_k_riccati_factor(self._h, m["N"], _dp(m["A"]), _dp(m["B"]), _dp(m["c"]),
                                                   _dp(m["Q"]), _dp(m["R"]), float(rho), _ip(m["block_type"]),
                                                   _dp(fac)))
        return fac

    def k_xupdate_riccati(self, N: int, fac: np.ndarray, has_c: bool, s0: np.ndarray, rt: np.ndarray):
        fac = np.ascontiguousarray(fac, dtype=np.float64)
        s0 = np.ascontiguousarray(s0, dtype=np.float64)
        rt = np.ascontiguousarray(rt, dtype=np.float64)
        x = np.empty_like(rt)
        self._check(self._L.admmb_k_xupdate_riccati(self._h, N, rt.shape[0], _dp(fac), int(has_c), _dp(s0),
                                                    _dp(rt), _dp(x)))
        return x

    def k_prox_dual_residuals(self, N, block_type, block_par, rinv, alpha, x, z, u):
        bt = np.ascontiguousarray(block_type, dtype=np.int32)
        bp = np.ascontiguousarray(block_par, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.array(z, dtype=np.float64, order="C")
        u = np.array(u, dtype=np.float64, order="C")
        rinv = np.ascontiguousarray(rinv, dtype=np.float64)
        norms = np.zeros((x.shape[0], 5))
        self._check(self._L.admmb_k_prox_dual_residuals(self._h, N, x.shape[0], _ip(bt), _dp(bp),
                                                        int(bp.shape[0] > 1), _dp(rinv), float(alpha), _dp(x),
                                                        _dp(z), _dp(u), _dp(norms)))
        return z, u, norms

    def k_dense_factor(self, N: int, fac: np.ndarray, has_c: bool):
        n = 9 * N + 6
        fac = np.ascontiguousarray(fac, dtype=np.float64)
        M, S, mc = np.zeros((n, n)), np.zeros((n, 6)), np.zeros(n)
        self._check(self._L.admmb_k_dense_factor(self._h, N, _dp(fac), int(has_c), _dp(M), _dp(S), _dp(mc)))
        return M, S, mc

    def k_xupdate_dense(self, N, M, S, mc, s0, rt, precision="fp64"):
        M, S, mc = (np.ascontiguousarray(a, dtype=np.float64) for a in (M, S, mc))
        s0 = np.ascontiguousarray(s0, dtype=np.float64)
        rt = np.ascontiguousarray(rt, dtype=np.float64)
        x = np.empty_like(rt)
        self._check(self._L.admmb_k_xupdate_dense(self._h, N, rt.shape[0], _dp(M), _dp(S), _dp(mc), _dp(s0),
                                                  _dp(rt), PRECISION[precision], _dp(x)))
        return x


def admm_solve(prob: dict, opts: dict, devices=None):
    """[x, z, u, hist] = admm_solve(prob, opts): same signature as oracle/admm_ocp.m."""
    with Solver(devices) as s:
        return s.solve(prob, opts)
