"""Small run through every kernel family, for compute-sanitizer (memcheck / racecheck) -- no oracle, few iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_pkg(); P = pkg.problems
with pkg.Solver() as s:
    for name, (prob, opts) in {
        "shared decoupled (cfg2)": P.cfg2_cw_batch(batch=70, N=7, seed=1),
        "shared decoupled N=1": P.cfg3_lowthrust_soc(batch=33, N=1, seed=1),
        "per-problem TMA (cfg4)": P.cfg4_elliptic(batch=45, N=6, seed=2),
        "per-problem TMA + affine + adaptive refactor": P.lqr_tracking(batch=40, N=5, seed=3, per_problem=True),
        "shared with cost, affine, q": P.lqr_tracking(batch=40, N=5, seed=4),
    }.items():
        o = dict(opts, max_iter=12, chunk=5, history=1)
        if "adaptive" in name:
            o.update(adapt_rho=1, adapt_every=3, adapt_mu=1.5)
        x, z, u, h = s.solve(prob, o)
        print(name, "ok", h["iters"][:3], h["launches"])
    prob, opts = P.cfg2_cw_batch(batch=50, N=6, seed=5)
    bt = prob["block_type"].copy(); bt[bt == P.BLK_NONE] = P.BLK_FREE
    x, z, u, h = s.solve(dict(prob, block_type=bt), dict(opts, max_iter=8))
    print("generic pattern ok")
    for prec in ("fp64", "tf32"):
        x, z, u, h = s.solve(prob, dict(opts, max_iter=6, xupdate="dense", precision=prec))
        print("dense", prec, "ok")
print("done")
