#!/bin/sh
# tools/collect_profiles.sh TAG ROUND - copy what profiles/README.md quotes from gpurun_out/TAG_* into profiles/rROUND_*
TAG=$1; R=$2
for f in bench_n1.json bench_ref.json launches.csv ncu_full_summary.csv rxloop.jsonl config1.json pytest.log smi.txt kernel_counters.csv chan.jsonl; do
  [ -f gpurun_out/${TAG}_$f ] && cp gpurun_out/${TAG}_$f profiles/r${R}_$f
done
cp gpurun_out/${TAG}_kernel_counters.json profiles/kernel_counters.json
ls -la profiles/
