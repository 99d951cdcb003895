"""Short, fixed-iteration run of the persistent ADMM kernel for ncu captures (no oracle, no torch).
usage: python scripts/profile_iterate.py [workload] [batch] [max_iter] [chunk]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 100
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 50
pkg = graft.load_pkg()
P = pkg.problems
gen = {"cfg2": P.cfg2_cw_batch, "cfg3": P.cfg3_lowthrust_soc, "cfg4": P.cfg4_elliptic, "cfg5": P.cfg5_montecarlo}[name]
prob, opts = gen(batch)
opts = dict(opts, max_iter=max_iter, chunk=chunk)
with pkg.Solver() as s:
    s.upload(prob, opts)
    for rep in range(2):
        t = time.time()
        r = s.run(opts)
        print(f"{name} batch={batch} iters={max_iter}: device {r['device_ms']:.3f} ms, kernel {r['kernel_ms']:.3f} ms "
              f"over {r['kernel_launches']} launches -> {r['stats'][1] / (r['kernel_ms'] * 1e-3):.4g} problem-iter/s "
              f"(wall {1e3 * (time.time() - t):.1f} ms)")
