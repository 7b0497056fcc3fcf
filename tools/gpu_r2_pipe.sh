#!/bin/sh
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/pipe.json 2>gpurun_out/pipe.err
tail -3 gpurun_out/pipe.err
python - <<'P'
import json
d=json.load(open('gpurun_out/pipe.json'))
w=d['wideband']
print('e2e', w['e2e']['value'], w['e2e']['ms_per_step'], 'pipelined', w['e2e_pipelined'])
print('resident', w['device_resident']['bursts_per_s'], w['payload_correct_frac'])
P
