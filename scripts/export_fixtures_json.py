"""Exports two committed golden fixtures as JSON for the Octave job (second half): arrays are stored
flattened in MATLAB's column-major order next to their MATLAB shape, so `reshape(data, shape)` rebuilds them.
  tests/golden/cfg1_single_n20.npz  -> tests/golden/octave/cfg1_single_n20.json   (admm_ocp.m)
  tests/golden/scp/scp_b6_n12.npz   -> tests/golden/octave/scp_b6_n12.json        (admm_scp.m)
Run here (CPU box): python scripts/export_fixtures_json.py.  The job that consumes them (.github/workflows/
oracle-octave.yml, oracle/check_octave.m) CANNOT be run in this image (no MATLAB / Octave): it is provided unexecuted."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

P = graft.load_pkg().problems
OUT = os.path.join(ROOT, "tests", "golden", "octave")
os.makedirs(OUT, exist_ok=True)


def mat(a, shape):
    """C-ordered array whose LAST axis is MATLAB's FIRST -> dict(shape=[MATLAB dims], data=[column-major values])."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return dict(shape=list(shape), data=[float(v) for v in a.ravel()])


def stage_mats(A):          # python (Bd, N, r, c) -> MATLAB [r, c, N, Bd]
    Bd, N, r, c = A.shape
    return mat(np.swapaxes(A, -1, -2), (r, c, N, Bd))


g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_single_n20.npz"))
N = int(g["in_A"].shape[1])
n, nb = 9 * N + 6, 3 * N + 2
doc = dict(
    what="config 1 (single CW impulsive rendezvous, N = 20): inputs, options and the oracle's outputs",
    prob=dict(A=stage_mats(g["in_A"]), B=stage_mats(g["in_B"]), s0=mat(g["in_s0"], (6, 1)),
              block_type=[int(v) for v in g["in_block_type"]], block_par=mat(g["in_block_par"], (8, nb, 1))),
    opts={k[4:]: float(g[k]) for k in g.files if k.startswith("opt_")},
    out=dict(x=mat(g["out_x"], (n, 1)), z=mat(g["out_z"], (n, 1)), u=mat(g["out_u"], (n, 1)),
             iters=[int(v) for v in g["out_iters"]], status=[int(v) for v in g["out_status"]]))
json.dump(doc, open(os.path.join(OUT, "cfg1_single_n20.json"), "w"))

g = np.load(os.path.join(ROOT, "tests", "golden", "scp", "scp_b6_n12.npz"))
B, N = int(g["batch"]), int(g["N"])
n, nb = 9 * N + 6, 3 * N + 2
prob, scp, opts = P.scp_nonlinear_rendezvous(B, N, seed=int(g["seed"]), scale=float(g["scale"]), substeps=int(g["substeps"]))
assert np.array_equal(prob["s0"], g["s0"])
doc = dict(
    what="SCP fixture (6 nonlinear-rendezvous problems, N = 12): inputs, options and the oracle's outputs",
    prob=dict(N=N, s0=mat(prob["s0"], (6, B)), block_type=[int(v) for v in prob["block_type"]],
              block_par=mat(prob["block_par"], (8, nb, 1))),
    opts={k: (v if isinstance(v, str) else float(v)) for k, v in opts.items()},
    scp={k: (v if isinstance(v, str) else float(v)) for k, v in scp.items()},
    out=dict(x=mat(g["x"], (n, B)), z=mat(g["z"], (n, B)), u=mat(g["u"], (n, B)),
             passes=[int(v) for v in g["passes"]], iters_total=[int(v) for v in g["iters_total"]],
             scp_status=[int(v) for v in g["scp_status"]]))
json.dump(doc, open(os.path.join(OUT, "scp_b6_n12.json"), "w"))
print(sorted(os.listdir(OUT)), [os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))])
