for pass in 1 2; do
for v in "" build/variants/libt9v0.so build/variants/libt9v1.so; do
  lib=""; [ -n "$v" ] && lib=$PWD/osmo_gmr_b200/$v
  echo "lib=$v"; GMR1B200_LIB=$lib python tools/bench_configs.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'])"
done; done
