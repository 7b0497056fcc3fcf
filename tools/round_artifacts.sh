#!/bin/sh
# tools/round_artifacts.sh TAG - on the GPU box: the whole GPU test suite, the bench line (with the CPU baseline),
# the reference arm, the ncu launch list of a bench run, one `ncu --set full` capture of each hot kernel and the
# receiver-loop bench.  Everything lands in gpurun_out/TAG_*; the files quoted in profiles/README.md are copied
# from there into profiles/ by hand.
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 600 gpurun_out/${TAG}_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"demod_kernel|decode_tpc|fcch" -s 10 -c 6 -f -o gpurun_out/${TAG}_full \
	python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for n in 1024 4096 16384; do python tools/bench_rxloop.py --channels $n; done > gpurun_out/${TAG}_rxloop.jsonl 2> gpurun_out/${TAG}_rxloop.err
ls -la gpurun_out/${TAG}_*
