#!/usr/bin/env bash
# Round-end evidence capture on one B200 (run via gpurun): GPU tests, the bench line with per-layer device times,
# the ncu launch list of one timed step and an `ncu --set full` pass over the 29 convolution launches of that step.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --layers > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final_layers.txt
head -c 400 gpurun_out/r01_bench_final.json; echo
L=50   # launches per step (gpu_launches / steps)
C=29   # convolution launches per step
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * L)) -c $L --csv --log-file gpurun_out/r01_launches.csv \
    $CMD > gpurun_out/r01_ncu_launches.log 2>&1
wc -l gpurun_out/r01_launches.csv
ncu --set full --clock-control none -k regex:conv_ -s $((3 * C)) -c $C -o /tmp/allconv $CMD > gpurun_out/r01_ncu_allconv.log 2>&1
ncu -i /tmp/allconv.ncu-rep --page raw --csv > gpurun_out/r01_allconv_raw.csv 2> /dev/null
ls -la gpurun_out/r01_allconv_raw.csv
