"""Generates tests/golden/scp/*.npz: a small seeded SCP batch solved by the oracle (oracle/scp_ocp.py
around oracle/admm_ocp_cpu.c) -- first-pass stage records, per-pass steps, passes, iteration totals and the final
x, z, u.  The reference tree has no SCP loop and no fixtures (README + LICENSE only); these vectors pin the oracle
(tests/test_scp.py, CPU) and the device loop is compared with them bit for bit (tests/test_scp.py, -m gpu).
Run here (CPU box):  python scripts/make_golden_scp.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import scp_ocp  # noqa: E402

P = graft.load_pkg().problems
OUT = os.path.join(ROOT, "tests", "golden", "scp")
os.makedirs(OUT, exist_ok=True)

for name, B, N, seed, scale, sub in (("scp_b6_n12", 6, 12, 61, 20.0, 3), ("scp_b5_n20_far", 5, 20, 62, 60.0, 4),
                                     ("scp_impulsive_b5_n12", 5, 12, 9, 30.0, 3), ("scp_elliptic_b5_n12", 5, 12, 64, 20.0, 3)):
    make = (P.scp_nonlinear_impulsive if "impulsive" in name else
            P.scp_nonlinear_elliptic if "elliptic" in name else P.scp_nonlinear_rendezvous)
    prob, scp, opts = make(B, N, seed=seed, scale=scale, substeps=sub)
    xref0, A0, B0, c0 = scp_ocp.shoot(prob["s0"], None, N, scp)
    x, z, u, info = scp_ocp.scp_solve(prob, scp, opts)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), batch=B, N=N, seed=seed, scale=scale, substeps=sub,
                        control=scp.get("control", "zoh"), model=scp.get("model", "nl_circular"),
                        s0=prob["s0"], xref0=xref0, A0=A0, B0=B0, c0=c0, x=x, z=z, u=u, passes=info["passes"],
                        scp_status=info["scp_status"], step=info["step"], iters_total=info["iters_total"],
                        hist_step=info["hist_step"], iters=info["iters"], status=info["status"])
    print(name, info["passes"], info["iters_total"])
