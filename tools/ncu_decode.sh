#!/bin/sh
# tools/ncu_decode.sh TAG [LIB ...] - instruction counts / issue utilisation / time of the first decode_tpc_kernel launches
# of a bench step under ncu (cold-cache, serialised: compare instruction counts and shares, not times), for the in-tree
# build and A/B builds.  smsp__inst_executed.sum (warp instructions) x 32 / (131 072 codewords x 3 392 state updates) is the
# "thread instructions per state update" figure of bench.py's viterbi block - re-measure it here whenever the trellis
# loop changes (the 7.8 of round 1's last build is a static SASS count on top of an ncu figure of 9.1).
TAG=$1; shift
mkdir -p gpurun_out
one() {
	GMR1B200_LIB=$2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum \
		--clock-control none -k regex:decode_tpc_kernel -s 2 -c 2 --csv --log-file gpurun_out/${TAG}_$1.csv \
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
    inst = float(v.get("smsp__inst_executed.sum", "0").replace(",", ""))
    print(sys.argv[1], k, "thread_instr_per_state_update=%.2f" % (inst * 32 / (131072 * 3392.0)),
          " ".join("%s=%s" % (a.split("__")[-1][:30], b) for a, b in v.items()))
PY
}
one main ""
for lib in "$@"; do one "$(basename $lib .so | sed 's/^lib//')" "$PWD/$lib"; done
