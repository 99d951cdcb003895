"""Per-launch trace (active problems, kernel ms) of one solve to tolerance:
ADMMB_TRACE=1 python scripts/trace_solve.py [batch] [chunk] [tf32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_pkg()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prob, opts = pkg.problems.cfg2_cw_batch(batch, 50, 1002)
if len(sys.argv) > 2 and int(sys.argv[2]) > 0:
    opts = dict(opts, chunk=int(sys.argv[2]))
if len(sys.argv) > 3 and sys.argv[3] == "tf32":
    opts = dict(opts, xupdate="dense", precision="tf32")
with pkg.Solver() as s:
    s.upload(prob, opts)
    import time
    s.run(opts)
    t0 = time.perf_counter()
    r = s.run(opts)
    print(r, "wall", time.perf_counter() - t0)
