#!/usr/bin/env bash
# Round-2 evidence capture on one B200 (run via gpurun, after the plain bench exited 0): `ncu --set full` over ALL launches of
# one timed 32 x 1080p step (convolutions: tensor-pipe activity; CBAM / glue: DRAM bytes), plus a source-level capture of the
# fused final dense block.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
L=${LAUNCHES_PER_STEP:-45}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none -s $((3 * L)) -c $L -o /tmp/${TAG}_step $CMD > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv 2> /dev/null
ls -la gpurun_out/${TAG}_step_raw.csv
ncu --set full --import-source on --clock-control none -k regex:dense_fused -s 3 -c 1 -o /tmp/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu_fused.log 2>&1
ncu -i /tmp/${TAG}_fused.ncu-rep --page source --csv > gpurun_out/${TAG}_fused_source.csv 2> /dev/null
ncu -i /tmp/${TAG}_fused.ncu-rep --page details --csv > gpurun_out/${TAG}_fused_details.csv 2> /dev/null
ls -la gpurun_out/${TAG}_fused_source.csv
