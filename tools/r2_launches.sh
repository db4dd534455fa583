#!/bin/bash
# launch lists (ncu gpu__time_duration) of one timed step per workload; usage: tools/r2_launches.sh TAG wl1 wl2 ...
TAG=$1; shift
mkdir -p gpurun_out/$TAG
for wl in "$@"; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-6000} -c 400 --csv --log-file gpurun_out/$TAG/launches_$wl.csv python bench.py --workload $wl --only --steps 2 --warmup 1 --no-cpu > gpurun_out/$TAG/ncu_$wl.log 2>&1
  python tools/launch_summary.py gpurun_out/$TAG/launches_$wl.csv > gpurun_out/$TAG/summary_$wl.txt 2>&1
  cat gpurun_out/$TAG/summary_$wl.txt
done
