#!/bin/sh
# two-block (2 x 4096) FCCH search vs the 8192-point form (GMR1B200_FCCH_FFT_SPLIT=0)
python -m pytest tests/test_sdr_gpu.py tests/test_chain_gpu.py tests/test_fullsize_gpu.py tests/test_rxsched_gpu.py tests/test_errors_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  GMR1B200_FCCH_FFT_SPLIT=$v python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-wideband --min-seconds 0 > gpurun_out/fft2_$v.json 2>gpurun_out/fft2_$v.err
  python - $v <<'P'
import json,sys
d=json.load(open('gpurun_out/fft2_%s.json'%sys.argv[1]))
print('split',sys.argv[1], round(d['value']/1e6,1), d['fcch']['ms_per_step'], d['fcch']['toa_identical_to_reference'] if 'toa_identical_to_reference' in d['fcch'] else None, d['configs']['4']['ms']['fcch_5_shift_search_and_fine'], round(d['configs']['4']['bursts_per_s']/1e6,1), [round(p['resident_bursts_per_s']/1e6,1) for p in d['sweep']['points']])
P
done
