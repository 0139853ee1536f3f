"""Time the 128x128 diagonal-block routine of the Cholesky panels in isolation (run via gpurun)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import build, _lib
build.build()
lib = _lib.load()
v = C.c_double()
for reps in (10, 50):
    assert lib.gpsat_microbench(0, 30, 0, reps, C.byref(v)) == 0, lib.gpsat_last_error()
    print(f"diag block 128x128 (potf2 + inverse + 5 tile products), {reps} reps: {v.value:.1f} us per block")
for mode, nm in ((0, "exp_neg vs exp(-u), u in [0, 745)"), (1, "sqrt_pos vs sqrt, x in [1e-36, 1e12)")):
    assert lib.gpsat_microbench(0, 40, mode, 2000, C.byref(v)) == 0, lib.gpsat_last_error()
    print(f"max relative error {nm}: {v.value:.3e}")
for kid, nm in enumerate(("Matern32", "Matern52", "Matern12", "RBF")):
    assert lib.gpsat_microbench(0, 21, kid, 2000, C.byref(v)) == 0
    print(f"kernel eval rate {nm}: {v.value:.1f} G entries/s")
for which, nm in ((5, "8 chains"), (6, "32 chains")):
    assert lib.gpsat_microbench(0, which, 0, 20000, C.byref(v)) == 0
    print(f"DMMA from one warp per scheduler, {nm}: {v.value:.2f} TFLOP/s")
