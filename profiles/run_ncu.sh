#!/bin/bash
# Run on the GPU box through gpurun: launch list + one full capture of the top kernel.
# usage: bash profiles/run_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-r01}
KREGEX=${2:-k_potrf_panel}
SKIP=${3:-40}
COUNT=${4:-3}
CMD="python bench.py --experts-per-step 592 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
