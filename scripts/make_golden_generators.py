"""Generates tests/golden/generators/*.npz: parameters + stage matrices of the spelled-out-order generator oracle
(oracle/gen_ocp.py) on small seeded cases.  The reference tree has no generator and no fixtures (README + LICENSE
only); these vectors pin the oracle generator (tests/test_oracle.py) and the device generators are compared with
them bit for bit (tests/test_gpu_units.py).   Run here (CPU box):  python scripts/make_golden_generators.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_ocp as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "generators")
os.makedirs(OUT, exist_ok=True)

rng = np.random.Generator(np.random.PCG64(31))
e = np.concatenate([[0.0, 0.05, 0.5, 0.7], rng.uniform(0.0, 0.8, 8)])
th = np.concatenate([[0.0, np.pi, 2 * np.pi, -1.0], rng.uniform(-10.0, 10.0, 8)])
N, T, sub = 6, 2.0 * np.pi / 6, 5
A, B = G.elliptic_stage_matrices(e, th, N, T, sub)
np.savez_compressed(os.path.join(OUT, "elliptic_b12_n6.npz"), kind="elliptic_zoh", e=e, theta0=th, N=N, T=T, substeps=sub,
                    A=A, B=B)
for kind, T, nmm in (("cw_impulsive", 2.0 * np.pi / 50, 1.0), ("cw_zoh", 2.0 * np.pi / 100, 1.0), ("cw_zoh", 0.37, 1.3)):
    A, B = G.cw_stage_matrices(kind, 4, T, nmm)
    np.savez_compressed(os.path.join(OUT, f"{kind}_T{T:.4f}_n{nmm}.npz"), kind=kind, N=4, T=T, nmm=nmm, A=A, B=B)
x = np.concatenate([np.linspace(-50.0, 50.0, 4001), rng.uniform(-1000.0, 1000.0, 2000)])
s, c = G.det_sincos(x)
np.savez_compressed(os.path.join(OUT, "det_sincos.npz"), x=x, s=s, c=c)
print(sorted(os.listdir(OUT)))
