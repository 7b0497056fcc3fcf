#!/bin/sh
# tools/ncu_demod.sh TAG [LIB ...] - instruction counts / occupancy / time of the first demod launches under ncu
# (cold-cache, serialised: compare instruction counts and shares, not times) for the in-tree build and A/B builds.
TAG=$1; shift
mkdir -p gpurun_out
one() {
	GMR1B200_LIB=$2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers \
		--clock-control none -k regex:demod_ -s 4 -c 2 --csv --log-file gpurun_out/${TAG}_$1.csv \
		python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
	python - "$1" gpurun_out/${TAG}_$1.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    out.setdefault(d["ID"], {})[d["Metric Name"]] = d["Metric Value"]
for k, v in out.items():
    print(sys.argv[1], k, " ".join("%s=%s" % (a.split("__")[-1][:28], b) for a, b in v.items()))
PY
}
one main ""
for lib in "$@"; do one "$(basename $lib .so | sed 's/^lib//')" "$PWD/$lib"; done
