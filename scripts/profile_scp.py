"""One short SCP solve for ncu: 4,096 nonlinear-rendezvous problems, two passes of at most 200 iterations.
usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:'k_scp_linearise|k_admm_iterate_pptma' \
                     -c 3 -o gpurun_out/r2_scp python scripts/profile_scp.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_pkg()
prob, scp, opts = pkg.problems.scp_nonlinear_rendezvous(4096, 50)
with pkg.Solver() as s:
    x, z, u, h = s.scp_solve(prob, dict(scp, max_pass=2), dict(opts, max_iter=200, chunk=100))
    print("passes", h["passes"].max(), "launches", h["launches"], "device ms", h["device_ms"])
