#!/bin/bash
# Slot-group count (GPSAT_GROUPS) x panel mode on the small-matrix workloads, c3 / c5 with more groups, and a
# compute-sanitizer memcheck pass over smoke().  Run on the GPU box through gpurun; logs in gpurun_out/gs_*.log
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --skip-e2e"
run() { name=$1; shift; env "$@" $B $ARGS > gpurun_out/gs_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/gs_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"])' 2>&1 | tail -1)"; }
ARGS="--workload c1 --experts-per-step 256 --steps 5 --warmup 3"
run c1_g3 X=1
run c1_g4 GPSAT_GROUPS=4
run c1_g6 GPSAT_GROUPS=6
run c1_g8 GPSAT_GROUPS=8
run c1_la_g8 GPSAT_GROUPS=8 GPSAT_PANEL_LA=6
ARGS="--workload c2 --steps 5 --warmup 3"
run c2_g3 X=1
run c2_g6 GPSAT_GROUPS=6
ARGS="--steps 2 --warmup 2"
run c3_g6 GPSAT_GROUPS=6
ARGS="--workload c5 --experts-per-step 256 --steps 2 --warmup 1"
run c5_g3 X=1
run c5_g6 GPSAT_GROUPS=6
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/gs_memcheck.log 2>&1
tail -5 gpurun_out/gs_memcheck.log
