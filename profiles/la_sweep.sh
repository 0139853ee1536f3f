#!/bin/bash
# Look-ahead panel mode (GPSAT_PANEL_LA) against the fused / safe modes and slot-group counts on the small-matrix
# workloads, c3 with the look-ahead forced on, and the stage stamps of the diagonal-block routine.
# Run on the GPU box through gpurun; logs land in gpurun_out/la_*.log
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --skip-e2e"
run() { name=$1; shift; env "$@" $B $ARGS > gpurun_out/la_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/la_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["phases"]["potrf"])' 2>&1 | tail -1)"; }
ARGS="--workload c1 --experts-per-step 256 --steps 5 --warmup 3"
run c1_la X=1
run c1_la0 GPSAT_PANEL_LA=0
run c1_safe GPSAT_PANEL_LA=0 GPSAT_SAFE_PANEL=1
run c1_la_g2 GPSAT_GROUPS=2
run c1_la_g4 GPSAT_GROUPS=4
run c1_la_g6 GPSAT_GROUPS=6
ARGS="--workload c2 --steps 5 --warmup 3"
run c2_la X=1
run c2_la0 GPSAT_PANEL_LA=0
ARGS="--steps 2 --warmup 2"
run c3_la99 GPSAT_PANEL_LA=99
python - <<'P' > gpurun_out/la_diag_stamps.log 2>&1
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from gpsat_b200 import _lib
lib = _lib.load()
def run(which, param, nk):
    v = C.c_double()
    rc = lib.gpsat_microbench(0, which, param, nk, C.byref(v))
    assert rc == 0, lib.gpsat_last_error()
    return v.value
print("us per 128x128 diagonal block:", run(30, 0, 50))
prev = 0.0
for st in range(1, 10):
    c = run(31, st, 20)
    print(f"stage {st}: {c:.0f} cycles from entry, +{c - prev:.0f}")
    prev = c
P
cat gpurun_out/la_diag_stamps.log
