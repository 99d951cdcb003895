"""Parallel-in-time kernel (opts.kernel = 'pint') against the C oracle on seeded cases: iteration counts and the
largest relative difference of x, z, u (its precision class: FP64, not the oracle's operation order).
usage (GPU box): python scripts/pint_check.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402
from oracle import cpu  # noqa: E402

pkg = graft.load_pkg()
P = pkg.problems
cpu.build()
cases = []
for N in (1, 2, 3, 7, 8, 9, 20, 50):
    prob, opts = P.cfg2_cw_batch(batch=97, N=N, seed=10 + N)
    cases.append((f"cfg2 N={N}", prob, dict(opts, max_iter=3000)))
prob, opts = P.cfg3_lowthrust_soc(batch=130, N=40, seed=5)
cases.append(("cfg3 N=40", prob, dict(opts, max_iter=3000)))
prob, opts = P.cfg5_montecarlo(batch=200, N=20, seed=6)
cases.append(("cfg5 adaptive N=20", prob, dict(opts, max_iter=4000, adapt_every=10, adapt_until=200, adapt_mu=5.0)))
prob, opts = P.cfg1_single_impulsive()
cases.append(("cfg1 history", prob, dict(opts, max_iter=3000, history=1)))
prob, opts = P.cfg2_cw_batch(batch=4500, N=50, seed=77)
cases.append(("cfg2 4500 x N=50 fixed 300 it", prob, dict(opts, max_iter=300)))
with pkg.Solver() as s:
    for name, prob, opts in cases:
        ref = cpu.solve(prob, opts)
        got = s.solve(prob, dict(opts, kernel="pint"))
        same = int((got[3]["iters"] == ref[3]["iters"]).sum())
        dmax = int(np.abs(got[3]["iters"].astype(np.int64) - ref[3]["iters"]).max())
        eq = got[3]["iters"] == ref[3]["iters"]
        rel = [float(np.abs(a[eq] - b[eq]).max() / max(1.0, np.abs(b[eq]).max())) if eq.any() else float("nan")
               for a, b in zip(got[:3], ref[:3])]
        print(f"{name:32s} iters equal {same}/{len(eq)} (max |diff| {dmax}), status equal "
              f"{bool(np.array_equal(got[3]['status'], ref[3]['status']))}, rel diff x {rel[0]:.2e} z {rel[1]:.2e} u {rel[2]:.2e}",
              flush=True)
