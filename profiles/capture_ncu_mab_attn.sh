set -u
TAG=r01r
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 -c $4 -f -o /tmp/${TAG}_$1 $BENCH > $OUT/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>> $OUT/${TAG}_$1.log
  echo "$1: $(grep -c . $OUT/${TAG}_$1_raw.csv) csv lines"
}
cap mab 'mab_(fwd|bwd)_kernel' 2 2
ncu -i /tmp/${TAG}_mab.ncu-rep --page source --csv > $OUT/${TAG}_mab_src.csv 2>/dev/null
cap attn 'attn_(fwd|bwd)_kernel' 6 2
