"""Print the actual relative errors of the CUDA path against the float64 oracle and the extended-precision fixture at
full size (diagnostic used to set the tolerances in tests/test_gpu_parity.py).  python profiles/diag_accuracy.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_extended import inputs
from oracle import gpr
from gpsat_b200 import get_engine
eng = get_engine(0)
sizes, Xs, zs, cs, theta, Xp = inputs()
ext = np.load(os.path.join(ROOT, "tests", "golden", "extended.npz"))
off = np.zeros(len(sizes) + 1, dtype=np.int64); off[1:] = np.cumsum(sizes)
b = eng.make_batch(off, np.concatenate(Xs), np.concatenate(zs), coords_scale=cs)
f, g = eng.eval(b, theta, grad=True)
P = len(Xp)
fm, fv, _, _ = eng.predict(b, theta, np.arange(len(sizes) + 1) * P, np.tile(Xp, (len(sizes), 1)))
f, g, fm, fv = f.cpu().numpy(), g.cpu().numpy(), fm.cpu().numpy(), fv.cpu().numpy()
for e, n in enumerate(sizes):
    fr, gr = gpr.neg_lml_and_grad(Xs[e] / cs, zs[e], theta[e, :3], theta[e, 3], theta[e, 4])
    m, v = ext[f"mean_{n}"], ext[f"fvar_{n}"]
    sl = slice(e * P, (e + 1) * P)
    print(f"N={n}: f vs ext {abs(f[e]-ext[f'f_{n}'])/abs(ext[f'f_{n}']):.2e}; grad vs oracle64 "
          f"{np.abs(g[e]-gr).max()/np.abs(gr).max():.2e} (max norm) {(np.abs(g[e]-gr)/np.abs(gr)).max():.2e} (elementwise); "
          f"mean vs ext {np.abs(fm[sl]-m).max()/np.abs(m).max():.2e} (max norm) {(np.abs(fm[sl]-m)/np.abs(m)).max():.2e} (elementwise); "
          f"var vs ext {np.abs(fv[sl]-v).max()/np.abs(v).max():.2e} (max norm) {(np.abs(fv[sl]-v)/np.abs(v)).max():.2e} (elementwise)")
