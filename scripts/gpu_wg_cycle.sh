#!/bin/bash
# GPU-box side of one tuning cycle of the warp-group kernel: parity subset, fixed-iteration rates, one ncu capture.
# usage (here): gpurun --timeout 900 -- 'bash scripts/gpu_wg_cycle.sh <tag> [widths]'
tag=${1:-wg}
widths=${2:-1024,4096,8192,65536}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wg" > gpurun_out/t_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_$tag.log
tail -4 gpurun_out/t_$tag.log
timeout 200 python scripts/variant_rates.py cfg2 200 100 $widths wg > gpurun_out/rates_$tag.log 2>&1
cat gpurun_out/rates_$tag.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_admm_iterate_wg -c 1 -o gpurun_out/$tag -f \
    python scripts/variant_rates.py cfg2 20 20 4096 wg > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
