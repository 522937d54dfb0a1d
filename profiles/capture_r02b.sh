#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): speaker-state family, round 2.  For `--model onlysp` and `--model sps`:
# (1) plain bench run (must exit 0), (2) launch list of one step with device times, (3) `ncu --set full` of the cell kernels and
# of the sequence-level attention kernels, raw pages exported as CSV.     usage: bash profiles/capture_r02b.sh <tag>
set -u
TAG=${1:-r02b}
OUT=gpurun_out
for M in onlysp sps; do
  BENCH="python bench.py --model $M --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
  $BENCH > $OUT/${TAG}_${M}_plain.json 2> $OUT/${TAG}_${M}_plain.err || { echo "plain run failed ($M)"; tail -5 $OUT/${TAG}_${M}_plain.err; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_${M}_launches.csv $BENCH > $OUT/${TAG}_${M}_launches.log 2>&1
  echo "$M launch list: $(grep -c . $OUT/${TAG}_${M}_launches.csv) lines"
  ncu --set full --clock-control none --import-source on -k "regex:sps_(fwd|bwd)_kernel" --launch-skip 6 -c 2 -f -o /tmp/${TAG}_${M}_cell $BENCH > $OUT/${TAG}_${M}_cell.log 2>&1
  ncu -i /tmp/${TAG}_${M}_cell.ncu-rep --page raw --csv > $OUT/${TAG}_${M}_cell_raw.csv 2>> $OUT/${TAG}_${M}_cell.log
  echo "$M cell: $(grep -c . $OUT/${TAG}_${M}_cell_raw.csv) raw lines"
done
BENCH="python bench.py --model onlysp --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k "regex:xattn_(fwd|bwd)_kernel" --launch-skip 12 -c 2 -f -o /tmp/${TAG}_xattn $BENCH > $OUT/${TAG}_xattn.log 2>&1
ncu -i /tmp/${TAG}_xattn.ncu-rep --page raw --csv > $OUT/${TAG}_xattn_raw.csv 2>> $OUT/${TAG}_xattn.log
echo "xattn: $(grep -c . $OUT/${TAG}_xattn_raw.csv) raw lines"
