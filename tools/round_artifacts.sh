#!/bin/sh
# tools/round_artifacts.sh TAG - on the GPU box (one GPU): the whole GPU test suite, the ncu counters of the build
# (profiles/kernel_counters.json, read by bench.py when the build id matches), the bench line (CPU baseline, configs 3/4,
# config-5 sweep), the reference arm, the ncu launch list of a bench run, one `ncu --set full` summary of the hot
# kernels, the receiver-loop bench and config 1.  Everything lands in gpurun_out/TAG_*; tools/collect_profiles.sh copies
# what profiles/README.md quotes into profiles/.
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python tools/kernel_counters.py --out gpurun_out/${TAG}_kernel_counters.json --csv gpurun_out/${TAG}_kernel_counters.csv
cp gpurun_out/${TAG}_kernel_counters.json profiles/kernel_counters.json
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 400 gpurun_out/${TAG}_bench_n1.json; tail -2 gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 300 gpurun_out/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"demod_|decode_tpc|fcch_" -s 18 -c 6 -f -o gpurun_out/${TAG}_full \
	python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --no-wideband --min-seconds 0 --streams 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pfb_|resamp_kernel" -s 4 -c 2 -f -o gpurun_out/${TAG}_full_chan \
	python tools/bench_chan.py --reps 2 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"fcch_fft_kernel<\(bool\)1|fcch_fft_kernel<true" -c 1 -f -o gpurun_out/${TAG}_full_grid \
	python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_chan.ncu-rep gpurun_out/${TAG}_full_grid.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.csv
rm -f gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_chan.ncu-rep gpurun_out/${TAG}_full_grid.ncu-rep
python tools/bench_chan.py > gpurun_out/${TAG}_chan.jsonl; python tools/bench_chan.py --chans 256 >> gpurun_out/${TAG}_chan.jsonl
for n in 1024 4096 16384; do python tools/bench_rxloop.py --channels $n; done > gpurun_out/${TAG}_rxloop.jsonl 2> gpurun_out/${TAG}_rxloop.err
python tools/bench_config1.py > gpurun_out/${TAG}_config1.json 2> gpurun_out/${TAG}_config1.err
ls -la gpurun_out/${TAG}_*
