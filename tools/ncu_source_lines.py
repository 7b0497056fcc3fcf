#!/usr/bin/env python3
"""tools/ncu_source_lines.py SOURCE.csv OBJ KERNEL_SUBSTR [PHASES.txt] - join ncu's SASS-level source page
(`ncu -i rep --page source --csv`) with the line table of the same build (`nvdisasm -g`): executed warp-instructions
and stall samples per source line, top lines first, and per stall reason.  OBJ must be the object the profiled
library was linked from."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, obj, sub = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
cub = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], check=True, capture_output=True, text=True).stdout.split("\n")
line_of, cur, inside = {}, None, False
for ln in txt:
    if ln.startswith(".text."):
        inside = sub in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ia, isrc, isam, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[2][ia], 16)
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
n_bursts = None
for r in rows[2:]:
    if len(r) < len(h):
        continue
    off = int(r[ia], 16) - base
    key, _ = line_of.get(off, (None, ""))
    s, e = int(r[isam]), int(r[iex])
    per[key][0] += s
    per[key][1] += e
    for i in stall_cols:
        if r[i] not in ("", "0"):
            per[key][2][h[i]] += int(r[i])
    tot_s += s
    tot_e += e
print("total samples", tot_s, "warp-instructions", tot_e)
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
print(f"{'line':28s} {'samples':>8s} {'%':>6s} {'inst/unit':>10s}  top stalls")
for key, (s, e, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:60]:
    name = f"{key[0]}:{key[1]}" if key else "?"
    tops = " ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
    print(f"{name:28s} {s:8d} {100.0 * s / tot_s:6.2f} {e / units:10.1f}  {tops}")
