#!/bin/bash
# c3 with look-ahead panels forced on at 4 / 6 slot groups; c5 and c4 (one GPU) at the library defaults.
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --skip-e2e"
run() { name=$1; shift; env "$@" $B $ARGS > gpurun_out/l2_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/l2_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["phases"]["potrf"])' 2>&1 | tail -1)"; }
ARGS="--steps 2 --warmup 2"
run c3_la99_g6 GPSAT_PANEL_LA=99 GPSAT_GROUPS=6
run c3_la99_g4 GPSAT_PANEL_LA=99 GPSAT_GROUPS=4
ARGS="--workload c5 --experts-per-step 256 --steps 2 --warmup 1"
run c5 X=1
ARGS="--workload c4 --experts-per-step 128 --steps 1 --warmup 1"
run c4 X=1
run c4_la99_g6 GPSAT_PANEL_LA=99 GPSAT_GROUPS=6
