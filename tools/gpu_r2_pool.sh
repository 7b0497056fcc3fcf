#!/bin/sh
# two wideband recordings in flight inside the full bench: stream-ordered pool with / without cross-stream reuse
for v in 0 1; do
  GMR1B200_POOL_INTERNAL_DEPS=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0 > gpurun_out/pool_$v.json 2>gpurun_out/pool_$v.err
  python - $v <<'P'
import json,sys
d=json.load(open('gpurun_out/pool_%s.json'%sys.argv[1]))
w=d['wideband']
print('internal_deps',sys.argv[1],'value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,2),d['e2e']['frac_of_ceiling'],'wide e2e',round(w['e2e']['value']/1e6,1),'pipelined',round(w['e2e_pipelined']['value']/1e6,1),w['e2e_pipelined']['ms_per_recording'],'int8',round(w['int8_recording']['e2e_pipelined']['value']/1e6,1))
P
done
