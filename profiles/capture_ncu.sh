#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): one `ncu --set full` capture per kernel family of a bench step, then exports the
# raw page as CSV (the .ncu-rep files exceed gpurun's 64 MiB return limit).   usage: bash profiles/capture_ncu.sh <tag>
# Precondition (B200_PROFILING.md): the same bench command has exited 0 without ncu first.
set -u
TAG=${1:-r01n}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
cap() {  # name, kernel regex, launch-skip, launch-count
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 -c $4 -f -o /tmp/${TAG}_$1 $BENCH > $OUT/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>> $OUT/${TAG}_$1.log
  echo "$1: $(grep -c . $OUT/${TAG}_$1_raw.csv) csv lines"
}
cap mab 'mab_(fwd|bwd)_kernel' 2 2
cap attn 'attn_(fwd|bwd)_kernel' 6 6
cap gemmw 'gemm3w_kernel' 29 29
cap gemm 'gemm3_kernel' 23 23
cap dln 'dln_(fwd|bwd)_kernel|colsum_partial' 12 12
