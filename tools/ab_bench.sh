#!/bin/sh
# tools/ab_bench.sh TAG [LIB ...] - run on the GPU box: bench.py (no CPU baseline) once for the in-tree library
# ("main") and once per extra library path (A/B builds made by tools/build_variant.sh), two passes each so that
# box-to-box and run-to-run noise can be told apart.  Writes gpurun_out/TAG_<name>.json and prints one summary
# line per run: name, M bursts/s, ms/step, demod ms/launch, HBM fraction, Viterbi ms/launch, FCCH ms/step.
TAG=$1; shift
mkdir -p gpurun_out
run() {
	name=$1; lib=$2
	GMR1B200_LIB=$lib python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
	python - "$name" gpurun_out/${TAG}_$name.json <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1], round(d["value"] / 1e6, 1), round(d["ms_per_step"], 4), round(r["ms_per_launch"], 4),
          round(r["frac"], 4), round(d["viterbi"]["ms_per_launch"], 4), round(d["fcch"]["ms_per_step"], 4),
          d["e2e"]["same_results_as_device_path"], d["crc_ok_frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
EOF
}
for pass in 1 2; do
	run main ""
	for lib in "$@"; do
		run "$(basename $lib .so | sed 's/^lib//')" "$PWD/$lib"
	done
done
