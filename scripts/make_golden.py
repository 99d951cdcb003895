"""Generates tests/golden/*.npz: inputs + outputs of the NumPy oracle (oracle/admm_ocp.py) on small seeded
cases.  The reference tree has no fixtures of its own (README + LICENSE only), so these vectors pin the
*oracle*: the C restatement and the CUDA path are both checked against them.
Run here (CPU box):  python scripts/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import admm_ocp as O  # noqa: E402

P = graft.load_pkg().problems
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def cases():
    prob, opts = P.cfg1_single_impulsive()
    yield "cfg1_single_n20", prob, dict(opts, max_iter=3000, history=1)
    prob, opts = P.cfg2_cw_batch(batch=8, N=10, seed=21)
    yield "cfg2_b8_n10", prob, dict(opts, max_iter=1500, history=1)
    prob, opts = P.cfg3_lowthrust_soc(batch=6, N=12, seed=22)
    yield "cfg3_b6_n12", prob, dict(opts, max_iter=600, history=1)
    prob, opts = P.cfg4_elliptic(batch=5, N=10, seed=23)
    yield "cfg4_b5_n10", prob, dict(opts, max_iter=600, history=1)
    prob, opts = P.cfg5_montecarlo(batch=8, N=10, seed=24)
    yield "cfg5_b8_n10_adaptive", prob, dict(opts, max_iter=1500, adapt_every=10, adapt_until=200, history=1)
    prob, opts = P.lqr_tracking(batch=5, N=8, seed=25)
    yield "lqr_b5_n8", prob, dict(opts, max_iter=200, history=1)
    prob, opts = P.lqr_tracking(batch=4, N=8, seed=26, per_problem=True)
    yield "lqr_pp_b4_n8_adaptive", prob, dict(opts, max_iter=200, adapt_rho=1, adapt_every=5, adapt_mu=2.0, history=1)


ARR = ("A", "B", "c", "Q", "R", "q", "s0", "block_type", "block_par")
for name, prob, opts in cases():
    x, z, u, h = O.admm_solve(prob, opts)
    d = {f"in_{k}": prob[k] for k in ARR if prob.get(k) is not None}
    d.update({f"opt_{k}": np.asarray(v) for k, v in opts.items() if not isinstance(v, str)})
    d.update(out_x=x, out_z=z, out_u=u, out_iters=h["iters"], out_status=h["status"], out_rho=h["rho"],
             out_r_norm=h["r_norm"], out_s_norm=h["s_norm"], out_hist_r=h["hist"]["r_norm"],
             out_hist_s=h["hist"]["s_norm"], out_refactor=np.asarray(h["refactor_count"]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "iters", h["iters"], "status", h["status"], "refactor", h["refactor_count"])
