"""Isolated measurements of the FP64 tensor pipe and the GEMM cores on this GPU (run via gpurun)."""
import ctypes as C
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import build, _lib
build.build()
lib = _lib.load()
out = {}
def run(which, param, nk):
    v = C.c_double()
    rc = lib.gpsat_microbench(0, which, param, nk, C.byref(v))
    assert rc == 0, lib.gpsat_last_error()
    return round(v.value, 2)
for which, nch in enumerate((1, 2, 4, 8)):
    for ctas in (1, 2, 4):
        out[f"dmma_chain acc/warp={nch} warps/SM={8*ctas}"] = run(which, ctas, 20000)
for mode, nm in ((0, "smem"), (1, "hbm"), (2, "l2")):
    out[f"core64 {nm}"] = run(10, mode, 256)
    out[f"core128 NT {nm}"] = run(11, mode, 256)
    out[f"core128 TN {nm}"] = run(12, mode, 256)
    out[f"core128 NN {nm}"] = run(13, mode, 256)
for nk in (2, 4, 8, 16):
    out[f"core128 NT hbm nk={nk}"] = run(11, 1, nk)
    out[f"core64 hbm nk={nk}"] = run(10, 1, nk)
for mode, nm in ((1, "dmma only"), (2, "dfma only"), (3, "dmma + dfma")):
    out[f"pipe mix {nm} (ms)"] = run(20, mode, 4000)
for kid, nm in enumerate(("Matern32", "Matern52", "Matern12", "RBF")):
    out[f"kernel eval rate {nm} (G entries/s)"] = run(21, kid, 2000)
out["diag block 128x128 potrf+inverse (us per block per CTA)"] = run(30, 0, 50)
print(json.dumps(out, indent=1))
