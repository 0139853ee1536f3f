"""One launch of the 128x128 core streaming 256 k-tiles per CTA (for ncu)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import build, _lib
build.build()
lib = _lib.load()
v = C.c_double()
assert lib.gpsat_microbench(0, 11, 1, 256, C.byref(v)) == 0
print("core128 NT hbm:", v.value)
assert lib.gpsat_microbench(0, 11, 0, 256, C.byref(v)) == 0
print("core128 NT smem:", v.value)
