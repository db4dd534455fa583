#!/bin/bash
# round 2: `ncu --set full` of the solver kernels, one launch each, after the settle steps (bench.py --only --steps 2).
# usage: bash tools/r2_ncu_full.sh TAG     -> gpurun_out/TAG/{tiles_chains,df_mixed,posdf_mixed}.ncu-rep
TAG=${1:-ncu}
mkdir -p gpurun_out/$TAG
run() {  # name kernel-regex workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${SKIP:-130} -c 1 -o gpurun_out/$TAG/$1 -f \
      python bench.py --workload $3 --only --steps 2 --warmup 1 --no-cpu > gpurun_out/$TAG/$1.log 2>&1
  ncu -i gpurun_out/$TAG/$1.ncu-rep --page details --csv > gpurun_out/$TAG/$1.details.csv 2>/dev/null
  ncu -i gpurun_out/$TAG/$1.ncu-rep --page raw --csv > gpurun_out/$TAG/$1.raw.csv 2>/dev/null
  python - "$TAG" "$1" <<'PY'
import csv, sys
tag, name = sys.argv[1:3]
rows = list(csv.reader(open(f"gpurun_out/{tag}/{name}.raw.csv")))
h, v = rows[0], rows[-1]
want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
for k in want:
    if k in h: print(f"{name}: {k} = {v[h.index(k)]} {rows[1][h.index(k)] if len(rows) > 2 else ''}")
PY
}
run tiles_chains k_island_tiles chains_1048576
run df_mixed k_solve_df mixed_262144
run posdf_mixed k_position_df mixed_262144
