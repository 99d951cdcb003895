#!/bin/bash
# GPU-box side of a whole-build check: the -m gpu suite, smoke(), the N=1 bench line, config-4 rates of the per-problem
# kernels and one ncu capture of the streamed-record warp-group kernel.
# usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round_check.sh <tag>'
tag=${1:-chk}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_$tag.log
tail -6 gpurun_out/t_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke exit $?"
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_$tag.json
timeout 300 python scripts/variant_rates.py cfg4 200 100 64,2048,4096,8192,12288,16384 auto,thread,wg > gpurun_out/rates_cfg4_$tag.log 2>&1
cat gpurun_out/rates_cfg4_$tag.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_admm_iterate_wg -c 1 -o gpurun_out/wgpp_$tag -f \
    python scripts/variant_rates.py cfg4 20 20 2048 wg > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
