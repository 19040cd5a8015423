#!/usr/bin/env bash
# Per-launch device times of one small bench run (B200_PROFILING.md: plain run first, then ncu launch list).
set -u
BATCH=${1:-4}
mkdir -p gpurun_out
CMD="python bench.py --batch $BATCH --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 156 -c 104 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
