"""SCP loop timings on the GPU box: python scripts/scp_rates.py [batch ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_pkg()
P = pkg.problems
sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096]
with pkg.Solver() as s:
    p0, s0, o0 = P.scp_nonlinear_rendezvous(64, 50)
    s.scp_solve(p0, dict(s0, max_pass=2), o0)
    for b in sizes:
        prob, scp, opts = P.scp_nonlinear_rendezvous(b, 50)
        t0 = time.perf_counter()
        x, z, u, h = s.scp_solve(prob, scp, opts)
        wall = 1e3 * (time.perf_counter() - t0)
        print(f"batch {b}: device {h['device_ms']:.1f} ms (wall {wall:.1f}), linearise {h['linearise_ms']:.2f} ms, "
              f"kernel {h['kernel_ms']:.1f} ms in {h['kernel_launches']} launches, passes {np.bincount(h['passes'])}, "
              f"converged {h['scp_stats'][0]}, ADMM iterations {h['scp_stats'][1]:.3e} -> "
              f"{h['scp_stats'][1] / h['device_ms'] / 1e-3:.3e} problem-iter/s, admm status {np.bincount(h['status'])}", flush=True)
