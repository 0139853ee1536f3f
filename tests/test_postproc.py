"""Post-processing: oracle vs the reference's golden outputs (CPU), CUDA path vs both (GPU).
Tolerance: the reference accumulates sequentially / by pandas group sums, the kernels by a fixed tree -> sums of
O(1e3..1e4) positive fp64 terms agree to ~1e-13 relative; asserted at 1e-11."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import postproc as opp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "postproc.npz"))
RTOL = 1e-11


def _wv_frame():
    return pd.DataFrame({"pred_loc_x": G["wv_px"], "pred_loc_y": G["wv_py"], "x": G["wv_ex"], "y": G["wv_ey"],
                         "f*": G["wv_f"], "f*_var": G["wv_fvar"]})


def _param_table(rng, n_t=3, with_dim=True):
    gx, gy = np.meshgrid(np.arange(-6, 7) * 200_000.0, np.arange(-6, 7) * 200_000.0)
    rows = []
    for t in range(n_t):
        keep = rng.random(gx.size) < 0.85
        for d in range(3 if with_dim else 1):
            f = pd.DataFrame({"x": gx.ravel()[keep], "y": gy.ravel()[keep], "t": 18326.0 + t})
            if with_dim:
                f["_dim_0"] = d
            f["lengthscales"] = np.exp(rng.normal(0, 1, keep.sum()))
            rows.append(f)
    df = pd.concat(rows, ignore_index=True)
    df = df.sample(frac=1.0, random_state=3).reset_index(drop=True)     # slices interleaved, like an appended table
    df.loc[rng.random(len(df)) < 0.07, "lengthscales"] = np.nan
    return df


# ---------------- CPU: oracle pinned to the reference ----------------
def test_oracle_gaussian_2d_weight_matches_reference():
    o = opp.gaussian_2d_weight(G["gw_x"], G["gw_y"], G["gw_x"], G["gw_y"], float(G["gw_lx"]), float(G["gw_ly"]),
                               G["gw_vals"])
    np.testing.assert_allclose(o, G["gw_out"], rtol=RTOL)
    assert np.all(np.isnan(G["gw_allnan"]))
    assert np.all(np.isnan(opp.gaussian_2d_weight(G["gw_x"][:5], G["gw_y"][:5], G["gw_x"][:7], G["gw_y"][:7], 1.0, 1.0,
                                                  np.full(7, np.nan))))


def test_oracle_weighted_values_and_glue_match_reference():
    df = _wv_frame()
    o = opp.get_weighted_values(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"],
                                float(G["wv_lengthscale"]))
    np.testing.assert_allclose(o.values, G["wv_out"], rtol=RTOL)
    g2 = opp.glue_local_predictions(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"],
                                    float(G["gl2_radius"]))
    np.testing.assert_allclose(g2.values, G["gl2_out"], rtol=RTOL)
    g1 = opp.glue_local_predictions(df, ["pred_loc_x"], ["x"], "f*", float(G["gl2_radius"]))
    np.testing.assert_allclose(g1.values, G["gl1_out"], rtol=RTOL)


# ---------------- GPU: CUDA path vs reference golden and oracle ----------------
@pytest.mark.gpu
def test_gpu_gaussian_2d_weight_golden():
    from gpsat_b200 import postprocessing as pp
    o = pp.gaussian_2d_weight(G["gw_x"], G["gw_y"], G["gw_x"], G["gw_y"], float(G["gw_lx"]), float(G["gw_ly"]),
                              G["gw_vals"])
    np.testing.assert_allclose(o, G["gw_out"], rtol=RTOL)
    o = pp.gaussian_2d_weight(G["gw_x"][:5], G["gw_y"][:5], G["gw_x"][:7], G["gw_y"][:7], 1.0, 1.0, np.full(7, np.nan))
    assert np.all(np.isnan(o))
    assert len(pp.gaussian_2d_weight(np.zeros(0), np.zeros(0), G["gw_x"], G["gw_y"], 1.0, 1.0, G["gw_vals"])) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("with_dim,vmin,vmax", [(True, None, None), (True, 0.2, 3.0), (False, None, 2.0)])
def test_gpu_smooth_table_matches_oracle(with_dim, vmin, vmax):
    from gpsat_b200 import postprocessing as pp
    df = _param_table(np.random.default_rng(5), with_dim=with_dim)
    ref = opp.smooth_table(df, "lengthscales", ["x", "y", "t"], 300_000.0, 250_000.0, max=vmax, min=vmin)
    out = pp.smooth_hyperparameter_table(df, "lengthscales", ["x", "y", "t"], 300_000.0, 250_000.0, max=vmax, min=vmin)
    assert list(out.columns) == list(ref.columns) and out.index.names == ref.index.names
    assert out.index.equals(ref.index)                                   # same rows, same order
    for c in ref.columns:
        np.testing.assert_allclose(out[c].values.astype(float), ref[c].values.astype(float), rtol=RTOL)


@pytest.mark.gpu
def test_gpu_weighted_values_and_glue_golden():
    from gpsat_b200 import postprocessing as pp
    df = _wv_frame()
    o = pp.get_weighted_values(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"],
                               lengthscale=float(G["wv_lengthscale"]))
    assert list(o.columns) == ["pred_loc_x", "pred_loc_y", "f*", "f*_var"]
    np.testing.assert_allclose(o.values, G["wv_out"], rtol=RTOL)
    full = pp.get_weighted_values(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], "f*", drop_weight_cols=False,
                                  lengthscale=float(G["wv_lengthscale"]))
    np.testing.assert_allclose(full["w_f*"].values / full["_w"].values, G["wv_out"][:, 2], rtol=RTOL)
    g2 = pp.glue_local_predictions_2d(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"],
                                      float(G["gl2_radius"]))
    np.testing.assert_allclose(g2.values, G["gl2_out"], rtol=RTOL)
    g1 = pp.glue_local_predictions_1d(df, "pred_loc_x", "x", "f*", float(G["gl2_radius"]))
    np.testing.assert_allclose(g1.values, G["gl1_out"], rtol=RTOL)
    with pytest.raises(AssertionError):
        pp.get_weighted_values(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], "f*")          # no lengthscale
    with pytest.raises(NotImplementedError):
        pp.get_weighted_values(df, ["pred_loc_x"], ["x"], "f*", weight_function="boxcar", lengthscale=1.0)


@pytest.mark.gpu
def test_gpu_glue_full_size_properties():
    """c1-sized frame (256 experts x ~5000 prediction points): weighted means lie inside the group's range and a
    constant field is reproduced exactly up to rounding."""
    from gpsat_b200 import postprocessing as pp
    rng = np.random.default_rng(1)
    n = 1_300_000
    px = rng.integers(-400, 401, n) * 5_000.0
    py = rng.integers(-400, 401, n) * 5_000.0
    ex = np.round(px / 200_000.0) * 200_000.0 + rng.integers(-1, 2, n) * 200_000.0
    ey = np.round(py / 200_000.0) * 200_000.0 + rng.integers(-1, 2, n) * 200_000.0
    v = rng.normal(size=n)
    df = pd.DataFrame({"pred_loc_x": px, "pred_loc_y": py, "x": ex, "y": ey, "f*": v, "one": np.full(n, 0.37)})
    o = pp.glue_local_predictions_2d(df, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "one"], 600_000.0)
    np.testing.assert_allclose(o["one"].values, 0.37, rtol=1e-13)
    g = df.groupby(["pred_loc_x", "pred_loc_y"])["f*"].agg(["min", "max"]).reset_index()
    assert len(g) == len(o)
    assert np.all(o["f*"].values >= g["min"].values - 1e-12) and np.all(o["f*"].values <= g["max"].values + 1e-12)
