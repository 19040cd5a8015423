#!/usr/bin/env bash
# ncu capture of tcgen05 conv launches inside a small bench run (B200_PROFILING.md recipe: plain run first).
# usage: tools/ncu_conv.sh <skip> <count> <out-name> [batch]
set -u
SKIP=${1:-112}; COUNT=${2:-3}; OUT=${3:-prof_conv}; BATCH=${4:-2}
mkdir -p gpurun_out
CMD="python bench.py --batch $BATCH --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s $SKIP -c $COUNT \
    -o gpurun_out/$OUT $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
