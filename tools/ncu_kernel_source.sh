#!/usr/bin/env bash
# Source-level ncu capture of ONE launch: tools/ncu_kernel_source.sh <kernel regex> <skip> <tag> [batch]
set -u
K=$1; SKIP=$2; TAG=$3; B=${4:-32}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:$K -s $SKIP -c 1 -o /tmp/$TAG python tools/fused_run.py $B > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2> /dev/null
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2> /dev/null
ls -la gpurun_out/${TAG}_source.csv
