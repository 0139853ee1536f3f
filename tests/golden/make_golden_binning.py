"""Golden vectors for the upstream binning from the REFERENCE's DataPrep.bin_data
(GPSat/dataprepper.py:230-407, scipy binned_statistic_2d inside), run unmodified under the module stubs of
make_golden.py.  Authoring container only:  python tests/golden/make_golden_binning.py"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg.install_stubs()
    from GPSat.dataprepper import DataPrep
    rng = np.random.default_rng(20200305)
    n = 20_000
    # along-track-like samples: many on bin edges (lattice multiples of the 50 km grid), some outside, some exactly
    # on the last edge
    x = rng.uniform(-1.2e6, 1.2e6, n)
    y = rng.uniform(-1.2e6, 1.2e6, n)
    snap = rng.random(n) < 0.2
    x[snap] = np.round(x[snap] / 50_000.0) * 50_000.0
    y[snap] = np.round(y[snap] / 50_000.0) * 50_000.0
    x[:50], y[50:100] = 1.0e6, 1.0e6            # on the rightmost edges
    x[100:120], y[120:140] = -1.0e6, -1.0e6      # on the leftmost edges
    z = rng.normal(0.3, 0.1, n)
    df = pd.DataFrame({"x": x, "y": y, "z": z})
    out = dict(x=x, y=y, z=z, x_range=np.array([-1.0e6, 1.0e6]), y_range=np.array([-1.0e6, 1.0e6]), grid_res=50_000.0)
    for st in ("mean", "count", "sum", "std", "min", "max"):
        b, (xc, yc) = DataPrep.bin_data(df, x_range=[-1.0e6, 1.0e6], y_range=[-1.0e6, 1.0e6], grid_res=50_000.0,
                                        x_col="x", y_col="y", val_col="z", bin_statistic=st)
        out[f"b2_{st}"] = b
    out.update(xc=xc, yc=yc)
    b1, xc1 = DataPrep.bin_data(df, x_range=[-1.0e6, 1.0e6], grid_res=12_500.0, x_col="x", val_col="z",
                                bin_statistic="mean", bin_2d=False)
    out.update(b1_mean=b1, xc1=xc1)
    np.savez_compressed(os.path.join(HERE, "binning.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
