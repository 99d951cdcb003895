"""Developer check on a GPU box: the warp-group kernel against the one-problem-per-thread kernel (both bit-exact to the oracle
by the parity suite) on a few shapes; prints where they differ."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_pkg(); P = pkg.problems
cases = [("cfg2 N=50", lambda: P.cfg2_cw_batch(96, 50, 7)), ("cfg2 N=100", lambda: P.cfg2_cw_batch(96, 100, 7)),
         ("cfg3 N=50", lambda: P.cfg3_lowthrust_soc(96, 50, 7)), ("cfg3 N=100", lambda: P.cfg3_lowthrust_soc(96, 100, 7)),
         ("cfg3 N=64", lambda: P.cfg3_lowthrust_soc(96, 64, 7)), ("cfg2 N=7", lambda: P.cfg2_cw_batch(33, 7, 7))]
for name, gen in cases:
    prob, opts = gen()
    for mi in (3, 50, 400):
        o = dict(opts, max_iter=mi)
        with pkg.Solver() as s:
            a = s.solve(prob, dict(o, kernel="wg"))
        with pkg.Solver() as s:
            b = s.solve(prob, dict(o, kernel="thread"))
        d = [float(np.nanmax(np.abs(a[i] - b[i]))) for i in range(3)]
        bad = np.where(np.abs(a[0] - b[0]).max(axis=1) > 0)[0]
        print(f"{name} max_iter={mi}: |dx| {d[0]:.3e} |dz| {d[1]:.3e} |du| {d[2]:.3e} iters equal {np.array_equal(a[3]['iters'], b[3]['iters'])} bad problems {bad[:12]}", flush=True)
