"""Which tolerance can each TF32 form reach?  Incremental (condensed, default) vs absolute (ADMMB_NO_CONDENSED=1)
against the FP64 Riccati path, on 256 CW problems.  usage: python scripts/tf32_tolerance_sweep.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_pkg()
prob, opts = pkg.problems.cfg2_cw_batch(256, 50, 11)
with pkg.Solver() as s:
    for tol in (1e-5, 1e-6, 1e-7, 1e-8):
        o = dict(opts, abstol=tol, reltol=tol, max_iter=40000)
        ref = s.solve(prob, o)[3]
        os.environ.pop("ADMMB_NO_CONDENSED", None)
        inc = s.solve(prob, dict(o, xupdate="dense", precision="tf32"))[3]
        os.environ["ADMMB_NO_CONDENSED"] = "1"
        ab = s.solve(prob, dict(o, xupdate="dense", precision="tf32"))[3]
        os.environ.pop("ADMMB_NO_CONDENSED", None)
        print(f"tol {tol:g}: converged fp64 {(ref['status'] == 0).sum()}  tf32-incremental {(inc['status'] == 0).sum()}  "
              f"tf32-absolute {(ab['status'] == 0).sum()} | mean iters fp64 {ref['iters'].mean():.0f} "
              f"incremental {inc['iters'].mean():.0f} absolute {ab['iters'].mean():.0f}")
