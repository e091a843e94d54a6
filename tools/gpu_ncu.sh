#!/bin/bash
# ncu --set full capture of the dominant kernel of each workload (one GPU, one launch each).
# The .ncu-rep files are large (gpurun_out is capped at 64 MiB), so the raw and source pages are
# exported to CSV on the box, gzip'd, and the report itself is kept only when small.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh <tag> wl1 wl2 ...'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in "$@"; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 3 -c 1 -f -o $OUT/prof_$wl \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > $OUT/ncu_$wl.log 2>&1
  tail -1 $OUT/ncu_$wl.log
  if [ -f $OUT/prof_$wl.ncu-rep ]; then
    ncu -i $OUT/prof_$wl.ncu-rep --page raw --csv > $OUT/prof_${wl}_raw.csv 2>/dev/null
    ncu -i $OUT/prof_$wl.ncu-rep --page details --csv > $OUT/prof_${wl}_details.csv 2>/dev/null
    ncu -i $OUT/prof_$wl.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/prof_${wl}_source.csv.gz
    sz=$(stat -c %s $OUT/prof_$wl.ncu-rep)
    echo "$wl report bytes: $sz"
    if [ "$sz" -gt 8000000 ]; then rm -f $OUT/prof_$wl.ncu-rep; fi
  fi
done
du -sh $OUT
