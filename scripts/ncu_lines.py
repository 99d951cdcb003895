"""Join an `ncu --page source --csv` (SASS view) export with `nvdisasm -g -c` line info of the same kernel and
aggregate stall samples / executed instructions per source line or per named line range.
usage: python scripts/ncu_lines.py <ncu_source.csv> <nvdisasm.sass> <kernel symbol substring> [file suffix] [ranges]
ranges: name:lo-hi,name:lo-hi,... (source lines of the file given by `file suffix`)"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, sym = sys.argv[1:4]
fsuffix = sys.argv[4] if len(sys.argv) > 4 else ""
ranges = []
if len(sys.argv) > 5:
    for item in sys.argv[5].split(","):
        name, r = item.split(":")
        lo, hi = r.split("-")
        ranges.append((name, int(lo), int(hi)))

# ---- nvdisasm: instruction index -> (file, line) for the wanted function
lines = []
cur = None
inside = False
for ln in open(sass):
    if ln.startswith(".text."):
        inside = sym in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        # with `nvdisasm -gi` an instruction is preceded by its whole inline chain, innermost first: the last
        # entry is the outermost call site (a line of the kernel body)
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
if len(body) != len(lines):
    print(f"warning: {len(body)} ncu instructions vs {len(lines)} nvdisasm instructions", file=sys.stderr)
per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = tot_i = 0
for r, loc in zip(body, lines):
    s = int(r[ci["# Samples"]] or 0)
    ie = int(r[ci["Instructions Executed"]] or 0)
    key = loc
    if ranges and loc and loc[0].endswith(fsuffix):
        for name, lo, hi in ranges:
            if lo <= loc[1] <= hi:
                key = ("range", name)
                break
    e = per_line[key]
    e[0] += s
    e[1] += ie
    for h in stall_cols:
        v = int(r[ci[h]] or 0)
        if v:
            e[2][h] += v
    tot_s += s
    tot_i += ie
print(f"total samples {tot_s}, instructions executed {tot_i}")
for key, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
    top = sorted(e[2].items(), key=lambda kv: -kv[1])[:4]
    tops = " ".join(f"{k[6:]}={v}" for k, v in top)
    name = f"{key[1]}" if key and key[0] == "range" else (f"{key[0].split('/')[-1]}:{key[1]}" if key else "?")
    print(f"{name:28s} samples {e[0]:7d} ({100.0 * e[0] / max(tot_s, 1):5.1f} %)  inst {e[1]:10d} ({100.0 * e[1] / max(tot_i, 1):5.1f} %)  {tops}")
