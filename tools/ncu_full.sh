#!/bin/sh
# tools/ncu_full.sh TAG [LIB ...] - one `ncu --set full` capture of the first timed BCCH demod launch per build
TAG=$1; shift
mkdir -p gpurun_out
one() {
	GMR1B200_LIB=$2 ncu --set full --import-source on --clock-control none -k regex:demod_ -s 4 -c 1 -f -o gpurun_out/${TAG}_$1 \
		python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
}
one main ""
for lib in "$@"; do one "$(basename $lib .so | sed 's/^lib//')" "$PWD/$lib"; done
ls -la gpurun_out/${TAG}_*
