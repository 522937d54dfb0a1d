#!/bin/bash
# Runs ON THE GPU BOX (via gpurun).  (1) plain bench run (must exit 0), (2) launch list of one step with device times,
# (3) one `ncu --set full` capture of the recurrence kernels, raw + source pages exported as CSV.
# usage: bash profiles/capture_r02.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
echo "launch list: $(grep -c . $OUT/${TAG}_launches.csv) lines"
ncu --set full --clock-control none --import-source on -k "regex:mab_(fwd|bwd)_kernel" --launch-skip 6 -c 2 -f -o /tmp/${TAG}_mab $BENCH > $OUT/${TAG}_mab.log 2>&1
ncu -i /tmp/${TAG}_mab.ncu-rep --page raw --csv > $OUT/${TAG}_mab_raw.csv 2>> $OUT/${TAG}_mab.log
ncu -i /tmp/${TAG}_mab.ncu-rep --page source --csv > $OUT/${TAG}_mab_source.csv 2>> $OUT/${TAG}_mab.log
echo "mab: $(grep -c . $OUT/${TAG}_mab_raw.csv) raw lines, $(grep -c . $OUT/${TAG}_mab_source.csv) source lines"
