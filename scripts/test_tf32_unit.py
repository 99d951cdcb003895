"""TF32 tensor-core dense x-update vs float64 reference (debug / accuracy report)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_pkg(); P = pkg.problems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = 9 * N + 6
rng = np.random.default_rng(0)
mode = sys.argv[2] if len(sys.argv) > 2 else "rand"
if mode == "ident":
    M = np.eye(n); S = np.zeros((n, 6)); mc = np.zeros(n)
else:
    M = rng.standard_normal((n, n)) / np.sqrt(n); S = rng.standard_normal((n, 6)); mc = rng.standard_normal(n)
for B in (200, 4096):
    s0 = rng.standard_normal((B, 6)); rt = rng.standard_normal((B, n))
    ref = rt @ M.T + s0 @ S.T + mc
    with pkg.Solver() as s:
        for prec in ("tf32_single", "tf32"):
            x = s.k_xupdate_dense(N, M, S, mc, s0, rt, prec)
            err = np.abs(x - ref)
            print(B, prec, "max abs err %.3e rel %.3e" % (err.max(), err.max() / np.abs(ref).max()),
                  "nonzero frac %.3f" % (x != 0).mean())
            if err.max() > 1e-2:
                print(" x[0,:6]  ", x[0, :6]); print(" ref[0,:6]", ref[0, :6])
                print(" x[1,:6]  ", x[1, :6]); print(" ref[1,:6]", ref[1, :6])
                bad = np.argwhere(err > 1e-2)
                print(" bad rows(problem) range", bad[:, 0].min(), bad[:, 0].max(), "cols(elem)", bad[:, 1].min(), bad[:, 1].max(), "count", len(bad))
