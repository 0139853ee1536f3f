#!/bin/bash
# Run on the GPU box through gpurun: launch list + one full capture of the top kernel.
# usage: bash profiles/run_ncu.sh <tag> <kernel-regex>
TAG=${1:-r01}
KREGEX=${2:-k_potrf_update}
CMD="python bench.py --experts-per-step 148 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1500 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 40 -c 3 \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
