#!/bin/bash
# GPU-box side: config-4 rates of the streamed-record warp-group kernel for developer builds with other ring depths
# (-DWG_PP_RING=n into lib_r<n>/) and the per-warp cycle timeline of one iteration (-DWG_TIMING build in lib_timing/)
tag=${1:-ring}
mkdir -p gpurun_out
for lib in lib lib_r4 lib_r16; do
    [ -f admm-library_b200/$lib/libadmm_b200.so ] || continue
    echo "== $lib" >> gpurun_out/rates_$tag.log
    ADMMB_LIB=admm-library_b200/$lib/libadmm_b200.so timeout 200 python scripts/variant_rates.py cfg4 200 100 64,2048,4096 wg >> gpurun_out/rates_$tag.log 2>&1
done
cat gpurun_out/rates_$tag.log
if [ -f admm-library_b200/lib_timing/libadmm_b200.so ]; then
    ADMMB_LIB=admm-library_b200/lib_timing/libadmm_b200.so timeout 100 python scripts/variant_rates.py cfg4 20 20 64 wg > gpurun_out/timing_raw_$tag.log 2>&1
    grep wgt gpurun_out/timing_raw_$tag.log | sort -k3,3n -k5,5n | uniq > gpurun_out/timeline_$tag.log
    grep wgk gpurun_out/timing_raw_$tag.log | sort -k2,2n | uniq > gpurun_out/stages_$tag.log
    cat gpurun_out/timeline_$tag.log | head -70
    head -60 gpurun_out/stages_$tag.log
fi
