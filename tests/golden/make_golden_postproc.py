"""Golden vectors for the post-processing functions, generated from the REFERENCE:

  * GPSat.postprocessing.gaussian_2d_weight (numba gufunc, unmodified)       -> postproc.npz  gw_*
  * GPSat.utils.get_weighted_values                                          -> postproc.npz  wv_*
  * GPSat.postprocessing.glue_local_predictions_2d / _1d                     -> postproc.npz  gl_*

Run in the authoring container only:  python tests/golden/make_golden_postproc.py
Besides the four stubs of make_golden.py, importing GPSat.postprocessing needs placeholder modules for
matplotlib / seaborn / dataclasses_json (plotting and config dataclasses, not used by these functions).
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def more_stubs():
    def mod(name, **a):
        m = types.ModuleType(name)
        for k, v in a.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")
    mod("matplotlib.backends")
    mod("matplotlib.backends.backend_pdf", PdfPages=object)
    mod("seaborn")

    def dataclass_json(*a, **k):
        return a[0] if a and isinstance(a[0], type) else (lambda c: c)
    mod("dataclasses_json", dataclass_json=dataclass_json, config=lambda **k: {})


def main():
    mg.install_stubs()
    more_stubs()
    import GPSat.postprocessing as pp
    from GPSat.utils import get_weighted_values
    rng = np.random.default_rng(20200305)
    out = {}
    # --- gaussian_2d_weight: an expert lattice (200 km spacing) with NaN holes, as smooth_hyperparameters calls it
    gx, gy = np.meshgrid(np.arange(-15, 16) * 200_000.0, np.arange(-15, 16) * 200_000.0)
    x, y = gx.ravel(), gy.ravel()
    keep = rng.random(len(x)) < 0.8
    x, y = x[keep], y[keep]
    vals = np.exp(rng.normal(0, 1, len(x)))
    vals[rng.random(len(x)) < 0.1] = np.nan
    out.update(gw_x=x, gw_y=y, gw_vals=vals, gw_lx=200_000.0, gw_ly=150_000.0,
               gw_out=pp.gaussian_2d_weight(x, y, x, y, 200_000.0, 150_000.0, vals))
    # all-NaN input -> NaN
    out["gw_allnan"] = pp.gaussian_2d_weight(x[:5], y[:5], x[:7], y[:7], 1.0, 1.0, np.full(7, np.nan))
    # --- get_weighted_values on an overlapping-prediction frame (prediction locations shared between experts)
    n = 20_000
    px = rng.integers(-40, 41, n) * 5_000.0
    py = rng.integers(-40, 41, n) * 5_000.0
    ex = np.round(px / 200_000.0) * 200_000.0 + rng.integers(-1, 2, n) * 200_000.0
    ey = np.round(py / 200_000.0) * 200_000.0 + rng.integers(-1, 2, n) * 200_000.0
    df = pd.DataFrame({"pred_loc_x": px, "pred_loc_y": py, "x": ex, "y": ey,
                       "f*": rng.normal(size=n), "f*_var": rng.random(n)})
    wv = get_weighted_values(df, ref_col=["pred_loc_x", "pred_loc_y"], dist_to_col=["x", "y"],
                             val_cols=["f*", "f*_var"], weight_function="gaussian", lengthscale=100_000.0)
    out.update(wv_px=px, wv_py=py, wv_ex=ex, wv_ey=ey, wv_f=df["f*"].values, wv_fvar=df["f*_var"].values,
               wv_lengthscale=100_000.0, wv_out=wv[["pred_loc_x", "pred_loc_y", "f*", "f*_var"]].values)
    # --- glue_local_predictions_2d / _1d
    gl = pp.glue_local_predictions_2d(df, pred_loc_cols=["pred_loc_x", "pred_loc_y"], xprt_loc_cols=["x", "y"],
                                      vars_to_glue=["f*", "f*_var"], inference_radius=400_000.0, R=3)
    out.update(gl2_out=gl[["pred_loc_x", "pred_loc_y", "f*", "f*_var"]].values, gl2_radius=400_000.0)
    gl1 = pp.glue_local_predictions_1d(df, pred_loc_col="pred_loc_x", xprt_loc_col="x", vars_to_glue="f*",
                                       inference_radius=400_000.0, R=3)
    out.update(gl1_out=gl1[["pred_loc_x", "f*"]].values)
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), **out)
    print("postproc.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
