#!/bin/bash
# Runs ON THE GPU BOX (via gpurun), final round-2 state: `ncu --set full` of the time-parallel kernel families of one ATV step
# (attention, both GEMM kernels, row-wise kernels); the recurrence kernels are captured by capture_r02.sh.  The raw pages are
# exported as CSV (the .ncu-rep files exceed gpurun's 64 MiB return limit).      usage: bash profiles/capture_r02d.sh r02d
# bench.py runs 3 warm-up + 1 timed + 2 event-timed steps: --launch-skip = 3 steps' worth of the family's launches.
set -u
TAG=${1:-r02d}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$BENCH > $OUT/${TAG}_plain2.json 2> $OUT/${TAG}_plain2.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain2.err; exit 1; }
cap() {  # name, kernel regex, launch-skip, launch-count
  ncu --set full --clock-control none -k "regex:$2" --launch-skip $3 -c $4 -f -o /tmp/${TAG}_$1 $BENCH > $OUT/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>> $OUT/${TAG}_$1.log
  echo "$1: $(grep -c . $OUT/${TAG}_$1_raw.csv) csv lines"
}
cap attn 'attn_(fwd|bwd)_kernel' 18 6
cap gemm 'gemm3_kernel' 90 30
cap gemmw 'gemm3w_kernel' 99 33
cap dln 'dln_(fwd|bwd)_kernel|colsum_partial' 72 24
