"""ctypes binding of include/admm_b200.h.  Fails loudly when the CUDA library is missing: there is
no CPU or PyTorch fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libadmm_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

OK, E_BADARG, E_CUDA, E_NCCL, E_NOMEM, E_NODEVICE, E_STATE = 0, -1, -2, -3, -4, -5, -6
ERROR_NAMES = {E_BADARG: "ADMMB_E_BADARG", E_CUDA: "ADMMB_E_CUDA", E_NCCL: "ADMMB_E_NCCL",
               E_NOMEM: "ADMMB_E_NOMEM", E_NODEVICE: "ADMMB_E_NODEVICE", E_STATE: "ADMMB_E_STATE"}

EXPORTS = ["admmb_version", "admmb_create", "admmb_destroy", "admmb_last_error", "admmb_device_count",
           "admmb_solve", "admmb_upload", "admmb_run", "admmb_download", "admmb_set_stream", "admmb_shift_resolve",
           "admmb_k_riccati_factor", "admmb_k_xupdate_riccati", "admmb_k_prox_dual_residuals",
           "admmb_k_dense_factor", "admmb_k_xupdate_dense",
           "admmb_upload_generated", "admmb_solve_generated", "admmb_k_generate", "admmb_nccl_gathers",
           "admmb_scp_solve", "admmb_k_scp_linearise"]


class Problem(C.Structure):
    _fields_ = [("N", C.c_int32), ("batch", C.c_int64), ("A", c_dp), ("dyn_batched", C.c_int32),
                ("B", c_dp), ("c", c_dp), ("Q", c_dp), ("R", c_dp), ("q", c_dp), ("q_batched", C.c_int32),
                ("s0", c_dp), ("block_type", c_ip), ("block_par", c_dp), ("par_batched", C.c_int32),
                ("z0", c_dp), ("u0", c_dp), ("rho0", c_dp)]


class Generator(C.Structure):
    _fields_ = [("kind", C.c_int32), ("substeps", C.c_int32), ("T", C.c_double), ("nmm", C.c_double),
                ("e", c_dp), ("theta0", c_dp)]


class Scp(C.Structure):
    _fields_ = [("model", C.c_int32), ("substeps", C.c_int32), ("T", C.c_double), ("nmm", C.c_double),
                ("R0", C.c_double), ("max_pass", C.c_int32), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("control", C.c_int32), ("e", c_dp), ("theta0", c_dp)]


class ScpResult(C.Structure):
    _fields_ = [("passes", c_ip), ("scp_status", c_ip), ("step", c_dp), ("iters_total", C.POINTER(C.c_int64)),
                ("hist_step", c_dp), ("stats", C.c_int64 * 4), ("linearise_ms", C.c_double)]


class Opts(C.Structure):
    _fields_ = [("rho", C.c_double), ("alpha", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
                ("max_iter", C.c_int32), ("adapt_rho", C.c_int32), ("adapt_mu", C.c_double),
                ("adapt_tau", C.c_double), ("adapt_every", C.c_int32), ("adapt_until", C.c_int32),
                ("xupdate", C.c_int32), ("precision", C.c_int32), ("history", C.c_int32), ("chunk", C.c_int32),
                ("kernel", C.c_int32), ("tf32_switch", C.c_int32), ("tf32_refresh", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("x", c_dp), ("z", c_dp), ("u", c_dp), ("iters", c_ip), ("status", c_ip),
                ("r_norm", c_dp), ("s_norm", c_dp), ("eps_pri", c_dp), ("eps_dual", c_dp), ("rho", c_dp),
                ("hist_r", c_dp), ("hist_s", c_dp), ("hist_eps_pri", c_dp), ("hist_eps_dual", c_dp),
                ("hist_rho", c_dp), ("stats", C.c_int64 * 4), ("device_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("launches", C.c_int64),
                ("kernel_ms", C.c_double), ("kernel_launches", C.c_int64)]


class AdmmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libadmm_b200.so.  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  admm-library_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.admmb_version.restype = C.c_int
    L.admmb_create.argtypes = [C.POINTER(H), C.POINTER(C.c_int), C.c_int]
    L.admmb_destroy.argtypes = [H]
    L.admmb_last_error.argtypes = [H]
    L.admmb_last_error.restype = C.c_char_p
    L.admmb_device_count.argtypes = [H]
    L.admmb_nccl_gathers.argtypes = [H]
    L.admmb_solve.argtypes = [H, C.POINTER(Problem), C.POINTER(Opts), C.POINTER(Result)]
    L.admmb_upload.argtypes = [H, C.POINTER(Problem), C.POINTER(Opts)]
    L.admmb_run.argtypes = [H, C.POINTER(Opts), C.POINTER(Result)]
    L.admmb_download.argtypes = [H, C.POINTER(Result)]
    L.admmb_set_stream.argtypes = [H, C.c_void_p]
    L.admmb_shift_resolve.argtypes = [H, C.c_int32, c_dp, C.POINTER(Opts), C.POINTER(Result)]
    L.admmb_k_riccati_factor.argtypes = [H, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_double, c_ip, c_dp]
    L.admmb_k_xupdate_riccati.argtypes = [H, C.c_int32, C.c_int64, c_dp, C.c_int32, c_dp, c_dp, c_dp]
    L.admmb_k_prox_dual_residuals.argtypes = [H, C.c_int32, C.c_int64, c_ip, c_dp, C.c_int32, c_dp,
                                              C.c_double, c_dp, c_dp, c_dp, c_dp]
    L.admmb_k_dense_factor.argtypes = [H, C.c_int32, c_dp, C.c_int32, c_dp, c_dp, c_dp]
    L.admmb_k_xupdate_dense.argtypes = [H, C.c_int32, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int32, c_dp]
    L.admmb_upload_generated.argtypes = [H, C.POINTER(Problem), C.POINTER(Generator), C.POINTER(Opts)]
    L.admmb_solve_generated.argtypes = [H, C.POINTER(Problem), C.POINTER(Generator), C.POINTER(Opts), C.POINTER(Result)]
    L.admmb_k_generate.argtypes = [H, C.c_int32, C.c_int64, C.POINTER(Generator), c_dp, c_dp]
    L.admmb_scp_solve.argtypes = [H, C.POINTER(Problem), C.POINTER(Scp), C.POINTER(Opts), C.POINTER(Result),
                                  C.POINTER(ScpResult)]
    L.admmb_k_scp_linearise.argtypes = [H, C.c_int32, C.c_int64, C.POINTER(Scp), C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp]
    for name in EXPORTS:
        if name not in ("admmb_last_error",):
            getattr(L, name).restype = C.c_int
    L.admmb_last_error.restype = C.c_char_p
    _lib = L
    return L
