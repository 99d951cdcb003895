"""Debug: generic-record TMA kernel with four warps per CTA (working sets wider than 32 x SMs) against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
from oracle import cpu
pkg = graft.load_pkg()
P = pkg.problems
with pkg.Solver() as s:
    for b in (4736, 4800, 6000):
        prob, opts = P.lqr_tracking(batch=b, N=12, seed=3, per_problem=True)
        opts = dict(opts, max_iter=300)
        x, z, u, h = s.solve(prob, opts)
        xo, zo, uo, ho = cpu.solve(prob, opts)
        print("lqr pp", b, "status", np.bincount(h["status"], minlength=3), "iters equal", np.array_equal(h["iters"], ho["iters"]),
              "x equal", np.array_equal(x, xo), "nan", np.isnan(x).sum(), flush=True)
    for b in (4736, 4768, 8192, 16384):
        prob, scp, opts = P.scp_nonlinear_rendezvous(b, 50)
        x, z, u, h = s.scp_solve(prob, dict(scp, max_pass=1), opts)
        print("scp pass 1", b, "status", np.bincount(h["status"], minlength=3), "nan x", np.isnan(x).sum(), "iters", h["iters"].min(), h["iters"].max(), flush=True)
