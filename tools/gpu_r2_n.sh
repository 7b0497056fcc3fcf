#!/bin/sh
# packed 16-bit Viterbi (two codewords per thread): parity, A/B against one codeword per thread (GMR1B200_DECODE_P16=0)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py tests/test_chain_gpu.py tests/test_rxsched_gpu.py tests/test_rxcall_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in p16 one; do
  if [ $v = one ]; then export GMR1B200_DECODE_P16=0; else unset GMR1B200_DECODE_P16; fi
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-sweep --min-seconds 0 > gpurun_out/n_bench_$v.json 2> gpurun_out/n_bench_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/n_bench_{v}.json").read().strip().splitlines()[-1])
    print(v, "value", round(d["value"] / 1e6, 1), "ms/step", round(d["ms_per_step"], 4), "serial", round(d["roofline"]["serial_ms_per_step"], 4), "viterbi ms", round(d["viterbi"]["ms_per_launch"], 4), "ACS/s", d["viterbi"]["acs_state_updates_per_s"] / 1e12, "same", d["e2e"]["same_results_as_device_path"], d["crc_ok_frac"])
    for c in ("3", "4"):
        print("  cfg", c, round(d["configs"][c]["bursts_per_s"] / 1e6, 1), d["configs"][c]["ms"], d["configs"][c].get("parity_vs_cpu_reference", {}).get("identical"))
except Exception as e:
    print(v, "FAILED", e)
PY
  tail -2 gpurun_out/n_bench_$v.err
done
