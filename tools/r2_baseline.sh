#!/bin/bash
# round-2 baseline: every BASELINE config at 1 GPU with the round-1 code, plus launch lists
set -x
mkdir -p gpurun_out/r2a
for wl in mixed_262144 chains_1048576 boxes_4096 spheres_65536; do
  timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu > gpurun_out/r2a/bench_$wl.json 2> gpurun_out/r2a/bench_$wl.err
  tail -c 2500 gpurun_out/r2a/bench_$wl.json
done
for wl in chains_1048576 boxes_4096 spheres_65536; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9650 -c 450 --csv --log-file gpurun_out/r2a/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/r2a/ncu_$wl.log 2>&1
  python tools/launch_summary.py gpurun_out/r2a/launches_$wl.csv > gpurun_out/r2a/summary_$wl.txt 2>&1
  cat gpurun_out/r2a/summary_$wl.txt
done
