"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into a small text file for profiles/:
key raw metrics per captured launch, the warp-stall breakdown, the SASS instruction mix per warp and the
hottest SASS lines.  usage: python scripts/ncu_summary.py <report.ncu-rep> <out.md> [warps] [iters]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
iters = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
buf = [f"# ncu summary of {rep}\n"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    buf.append(f"\n## launch: {name[:120]}\n")
    for i, h in enumerate(hdr):
        if h in KEYS:
            buf.append(f"- {h}: {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r != hdr]
    # a multi-launch report repeats the table per launch: keep the first one
    if "Address" in hdr:
        seen, first = set(), []
        for r in data:
            a = r[hdr.index("Address")]
            if a in seen:
                break
            seen.add(a)
            first.append(r)
        data = first
    ci = {h: i for i, h in enumerate(hdr)}
    stall = collections.Counter()
    for r in data:
        for h in hdr:
            if h.startswith("stall_"):
                try:
                    stall[h] += int(r[ci[h]])
                except ValueError:
                    pass
    S = sum(stall.values()) or 1
    buf.append("\n## warp stall sampling (first captured launch)\n")
    for h, v in stall.most_common(12):
        buf.append(f"- {h}: {100 * v / S:.1f}%")
    ops = collections.Counter()
    for r in data:
        t = r[ci["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] += int(r[ci["Instructions Executed"]] or 0)
    T = sum(ops.values())
    buf.append(f"\n## SASS instruction mix: {T} warp-instructions"
               + (f" = {T / warps / iters:.0f} per warp per iteration" if warps * iters > 1 else "") + "\n")
    for op, v in ops.most_common(18):
        buf.append(f"- {op}: {100 * v / T:.1f}%" + (f" ({v / warps / iters:.0f}/warp-iter)" if warps * iters > 1 else ""))
    buf.append("\n## hottest SASS lines by samples\n")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:15]:
        buf.append(f"- {r[ci['# Samples']]:>8s}  {r[ci['Source']].strip()[:100]}")
open(out, "w").write("\n".join(buf) + "\n")
print("\n".join(buf[:40]))
