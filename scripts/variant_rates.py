"""Fixed-iteration throughput of each kernel variant of the FP64 Riccati path over a sweep of working-set widths
(no early exit: tolerances are zero, so every problem runs max_iter iterations).  No oracle, no torch.
usage: python scripts/variant_rates.py [workload] [max_iter] [chunk] [widths,comma] [variants,comma]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
widths = [int(w) for w in (sys.argv[4] if len(sys.argv) > 4 else "1024,4096,8192,16384,65536").split(",")]
variants = (sys.argv[5] if len(sys.argv) > 5 else "thread_wide,thread2,tile,wg").split(",")
pkg = graft.load_pkg()
if os.environ.get("ADMMB_LIB"):          # developer build of the library (e.g. -DWG_TIMING in lib_timing/)
    pkg._lib.LIB_PATH = os.path.abspath(os.environ["ADMMB_LIB"])
P = pkg.problems
gen = {"cfg2": P.cfg2_cw_batch, "cfg3": P.cfg3_lowthrust_soc, "cfg4": P.cfg4_elliptic, "cfg5": P.cfg5_montecarlo}[name]
for w in widths:
    prob, opts = gen(w)
    for v in variants:
        o = dict(opts, max_iter=max_iter, chunk=chunk, abstol=0.0, reltol=0.0, kernel=v)
        with pkg.Solver() as s:
            s.upload(prob, o)
            best = None
            for rep in range(3):
                r = s.run(o)
                if best is None or r["kernel_ms"] < best["kernel_ms"]:
                    best = r
        rate = best["stats"][1] / (best["kernel_ms"] * 1e-3)
        print(f"{name} width={w:7d} kernel={v:12s} {best['kernel_ms'] / max_iter * 1e3:9.2f} us/iteration  "
              f"{rate:.4g} problem-iter/s  ({best['kernel_launches']} launches)", flush=True)
