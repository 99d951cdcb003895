#!/bin/bash
# GPU-box side: parity subset + fixed-iteration rates of the warp-group kernel, then (second library, built with
# -DWG_TIMING into lib_timing/) the per-warp cycle timeline of one iteration
tag=${1:-wg}
widths=${2:-1024,4096,8192,65536}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wg" > gpurun_out/t_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_$tag.log
tail -3 gpurun_out/t_$tag.log
timeout 200 python scripts/variant_rates.py cfg2 200 100 $widths wg > gpurun_out/rates_$tag.log 2>&1
cat gpurun_out/rates_$tag.log
if [ -f admm-library_b200/lib_timing/libadmm_b200.so ]; then
    ADMMB_LIB=admm-library_b200/lib_timing/libadmm_b200.so timeout 100 python scripts/variant_rates.py cfg2 20 20 4096 wg 2>&1 | grep wgt | sort -k3,3n -k5,5n | uniq > gpurun_out/timeline_$tag.log
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
if [ -f admm-library_b200/lib_timing/libadmm_b200.so ]; then
    ADMMB_LIB=admm-library_b200/lib_timing/libadmm_b200.so timeout 100 python scripts/variant_rates.py cfg2 20 20 4096 wg 2>&1 | grep wgk | sort -k2,2n | uniq > gpurun_out/stages_$tag.log
    awk '{print $2, $4, $6, $8}' gpurun_out/stages_$tag.log | sort -n | uniq | awk 'NR%1==0' | head -60
fi
