"""128x64 core with two CTAs per SM against the 128x128 core, on streams of short GEMM tasks (run via gpurun)."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import build, _lib
build.build()
lib = _lib.load()
v = C.c_double()
out = {}
for nk in (1, 2, 4, 8, 16, 32):
    for which, nm in ((51, "128x128 x1/SM"), (50, "128x64 x2/SM")):
        rc = lib.gpsat_microbench(0, which, 16, nk, C.byref(v))
        out[f"{nm} nk={nk} (16 tasks/CTA)"] = round(v.value, 2) if rc == 0 else lib.gpsat_last_error().decode()
print(json.dumps(out, indent=1))
