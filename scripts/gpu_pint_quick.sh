#!/bin/bash
# GPU-box side: parallel-in-time kernel against the oracle, fixed-iteration rates, per-warp timeline (lib_timing build)
tag=${1:-pint}
widths=${2:-1024,4096,65536}
mkdir -p gpurun_out
timeout 600 python scripts/pint_check.py > gpurun_out/check_$tag.log 2>&1
tail -14 gpurun_out/check_$tag.log
timeout 200 python scripts/variant_rates.py cfg2 200 100 $widths wg,pint > gpurun_out/rates_$tag.log 2>&1
cat gpurun_out/rates_$tag.log
if [ -f admm-library_b200/lib_timing/libadmm_b200.so ]; then
    ADMMB_LIB=admm-library_b200/lib_timing/libadmm_b200.so timeout 100 python scripts/variant_rates.py cfg2 20 20 4096 pint 2>&1 | grep ptt | sort -k3,3n -k5,5n | uniq > gpurun_out/timeline_$tag.log
    python - <<PY
import collections
rows = collections.defaultdict(dict)
for ln in open("gpurun_out/timeline_$tag.log"):
    f = ln.split()
    rows[int(f[2])].setdefault(f[3], int(f[4]))
for w in sorted(rows):
    print("warp", w, " ".join(f"{k}={v}" for k, v in sorted(rows[w].items(), key=lambda kv: kv[1])))
PY
fi
