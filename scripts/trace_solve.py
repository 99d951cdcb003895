"""Per-launch trace (active problems, kernel ms) of one solve to tolerance:
ADMMB_TRACE=1 python scripts/trace_solve.py [batch] [chunk]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_pkg()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prob, opts = pkg.problems.cfg2_cw_batch(batch, 50, 1002)
if len(sys.argv) > 2:
    opts = dict(opts, chunk=int(sys.argv[2]))
with pkg.Solver() as s:
    s.upload(prob, opts)
    s.run(opts)
    r = s.run(opts)
    print(r)
