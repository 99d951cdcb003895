"""Short run of the dense shared-factor path (TF32 tensor-core GEMM + streaming prox kernel) for ncu captures.
usage: python scripts/profile_dense.py [batch] [max_iter] [precision]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_pkg()
if os.environ.get("ADMMB_LIB"):          # developer build of the library
    pkg._lib.LIB_PATH = os.path.abspath(os.environ["ADMMB_LIB"])
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 40
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
prob, opts = pkg.problems.cfg2_cw_batch(batch, 50, 2)
opts = dict(opts, max_iter=max_iter, xupdate="dense", precision=prec, chunk=max_iter)
with pkg.Solver() as s:
    s.upload(prob, opts)
    for rep in range(2):
        r = s.run(opts)
        print(f"dense {prec} batch={batch} iters={max_iter}: device {r['device_ms']:.3f} ms, GEMM+prox {r['kernel_ms']:.3f} ms "
              f"-> {r['stats'][1] / (r['kernel_ms'] * 1e-3):.4g} problem-iter/s")
