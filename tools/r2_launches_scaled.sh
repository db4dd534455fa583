#!/bin/bash
# launch list of one step of chains at 1/8 size (what one rank of an 8-GPU run steps): where the fixed per-step cost sits
TAG=$1; SCALE=${2:-0.125}
mkdir -p gpurun_out/$TAG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-6000} -c 400 --csv --log-file gpurun_out/$TAG/launches_chains_x$SCALE.csv python bench.py --workload chains_1048576 --scale $SCALE --only --steps 2 --warmup 1 --no-cpu > gpurun_out/$TAG/ncu_chains_x$SCALE.log 2>&1
python tools/launch_summary.py gpurun_out/$TAG/launches_chains_x$SCALE.csv > gpurun_out/$TAG/summary_chains_x$SCALE.txt 2>&1
cat gpurun_out/$TAG/summary_chains_x$SCALE.txt
