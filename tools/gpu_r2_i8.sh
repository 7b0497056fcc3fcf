#!/bin/sh
python -m pytest tests/test_chan_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/i8.json 2>gpurun_out/i8.err
tail -3 gpurun_out/i8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/i8.json'))
w=d['wideband']
print('int16 e2e', w['e2e']['value'], 'pipelined', w['e2e_pipelined']['value'], w['payload_correct_frac'])
w8=w['int8_recording']
print('int8 e2e', w8['e2e']['value'], w8['e2e']['ms_per_step'], 'pipelined', w8['e2e_pipelined']['value'], w8['e2e_pipelined']['ms_per_recording'], 'resident', w8['device_resident']['bursts_per_s'], w8['payload_correct_frac'], w8['crc_ok_but_payload_wrong'], w8['fcch_found_frac'], w8['int8_peak'])
P
