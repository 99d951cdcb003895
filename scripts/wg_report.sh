#!/bin/bash
# here: per-role aggregation of gpurun_out/<tag>.ncu-rep (needs the matching lib/obj/iter_wg.o)
tag=${1:-wg}
ranges=$2
mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/admm-library_b200/lib/obj/iter_wg.o >/dev/null 2>&1
nvdisasm -gi -c iter_wg.sm_100a.cubin > /tmp/cub/$tag.sass
cd /root/repo
ncu -i gpurun_out/$tag.ncu-rep --page source --csv 2>/dev/null > /tmp/${tag}_src.csv
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv 2>/dev/null > /tmp/${tag}_raw.csv
python scripts/ncu_lines.py /tmp/${tag}_src.csv /tmp/cub/$tag.sass "wgILb0ELb1" iterate_wg.cuh "$ranges" | head -16
python - <<PY
import csv
rows = list(csv.reader(open('/tmp/${tag}_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","sm__cycles_elapsed.max","launch__grid_size"]
for h,u,v in zip(hdr,units,vals):
    if h in want or any(h == w + ".pct_of_peak_sustained_elapsed" for w in want): print(h,u,v)
PY
