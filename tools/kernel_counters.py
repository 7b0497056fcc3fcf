#!/usr/bin/env python3
"""tools/kernel_counters.py [--out profiles/kernel_counters.json] - on the GPU box: ncu counters of the hot kernels of
one bench.py step (config 2: 131 072 BCCH + 131 072 DC6 bursts, 1 024 FCCH windows), tagged with the build id of the
library they were taken on.  bench.py reads the file for roofline.traffic (DRAM bytes per demod launch) and for the
instructions per Viterbi state update, and ignores it when the build id differs from the library it is timing.

ncu replays kernels serialised and cold-cache: the COUNTERS (bytes, instructions) are what is used, never its times."""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "kernel_counters.json"))
ap.add_argument("--csv", default=os.path.join(ROOT, "gpurun_out", "kernel_counters.csv"))
args = ap.parse_args()
os.makedirs(os.path.dirname(args.csv), exist_ok=True)
metrics = "smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum," \
          "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
cmd = ["ncu", "--metrics", metrics, "--clock-control", "none", "-k", "regex:demod_|decode_tpc|fcch_", "-s", "18", "-c", "6",
       "--csv", "--log-file", args.csv, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3",
       "--no-cpu-baseline", "--no-configs", "--no-sweep", "--no-wideband", "--min-seconds", "0", "--streams", "1"]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
rows = [r for r in csv.reader(open(args.csv)) if len(r) > 10]
h = rows[0]
per = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    per.setdefault((d["ID"], d["Kernel Name"]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
import osmo_gmr_b200
build = osmo_gmr_b200.Lib().version().split("build ")[-1]
units = {"demod_fast_kernel<0,81,0>": ("demod_bcch", 131072), "demod_fast_kernel<2,41,0>": ("demod_dc6", 131072),
         "decode_tpc_kernel<0,": ("decode_bcch", 131072), "decode_tpc_kernel<1,": ("decode_ccch", 131072),
         "fcch_fft_kernel<0>": ("fcch_rough", 1024), "fcch_grid_kernel<0>": ("fcch_rough", 1024),
         "fcch_rough_kernel": ("fcch_rough", 1024),
         "fcch_fine_kernel": ("fcch_fine", 1024)}
out = {}
for (kid, name), m in per.items():
    for pat, (key, n) in units.items():
        if pat in name.replace(" ", "").replace("(int)", "").replace("(bool)", "").replace("gmr1::", ""):
            out[key] = {"kernel": name, "units": n, "warp_inst": m["smsp__inst_executed.sum"],
                        "warp_inst_per_unit": m["smsp__inst_executed.sum"] / n,
                        "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
                        "dram_bytes_read": m["dram__bytes_read.sum"], "ncu_time_us": m["gpu__time_duration.sum"] / 1e3,
                        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                        "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"]}
json.dump({"build": build, "how": " ".join(cmd[:14]) + " python bench.py --steps 1 --warmup 3 --streams 1 ...",
           "kernels": out}, open(args.out, "w"), indent=1)
print(json.dumps({"build": build, "kernels": {k: round(v["warp_inst_per_unit"], 1) for k, v in out.items()}}))
