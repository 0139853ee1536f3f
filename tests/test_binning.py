"""Upstream binning: oracle vs the reference's golden outputs (CPU); CUDA path vs both (GPU).
Bin membership (counts) must be exact; sums are fp64 atomics (order differs from np.bincount) -> 1e-12 relative."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import binning as ob

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "binning.npz"))
XR, YR, RES = list(G["x_range"]), list(G["y_range"]), float(G["grid_res"])


def _frame():
    return pd.DataFrame({"x": G["x"], "y": G["y"], "z": G["z"]})


def test_oracle_bin_data_matches_reference():
    df = _frame()
    for st in ("mean", "count", "sum", "std", "min", "max"):
        b, (xc, yc) = ob.bin_data(df, XR, YR, RES, val_col="z", bin_statistic=st)
        np.testing.assert_array_equal(b, G[f"b2_{st}"])
    np.testing.assert_array_equal(xc, G["xc"])
    b1, xc1 = ob.bin_data(df, XR, YR, 12_500.0, val_col="z", bin_2d=False)
    np.testing.assert_array_equal(b1, G["b1_mean"])


@pytest.mark.gpu
def test_gpu_bin_data_golden():
    from gpsat_b200.dataprepper import DataPrep
    df = _frame()
    b, (xc, yc) = DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic="count")
    np.testing.assert_array_equal(b, G["b2_count"])                      # membership incl. edge rules: exact
    np.testing.assert_array_equal(xc, G["xc"])
    np.testing.assert_array_equal(yc, G["yc"])
    b, _ = DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic="mean")
    assert np.array_equal(np.isnan(b), np.isnan(G["b2_mean"]))
    np.testing.assert_allclose(b, G["b2_mean"], rtol=1e-12)
    b, _ = DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic="sum")
    np.testing.assert_allclose(b, G["b2_sum"], rtol=1e-12, atol=1e-13)
    # second-pass statistics (examples/bin_data.py:165 bins ["mean", "std", "count"]): extrema exact, std to rounding
    for st in ("min", "max"):
        b, _ = DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic=st)
        np.testing.assert_array_equal(b, G[f"b2_{st}"])                  # NaN in the same (empty) bins, same values
    b, _ = DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic="std")
    assert np.array_equal(np.isnan(b), np.isnan(G["b2_std"]))
    np.testing.assert_allclose(b, G["b2_std"], rtol=1e-11, atol=1e-15)   # single-sample bins: exactly 0 in both
    b1, xc1 = DataPrep.bin_data(df, x_range=XR, grid_res=12_500.0, x_col="x", val_col="z", bin_2d=False)
    np.testing.assert_array_equal(xc1, G["xc1"])
    np.testing.assert_allclose(b1, G["b1_mean"], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=RES, val_col="z", bin_statistic="median")
    with pytest.raises(AssertionError):
        DataPrep.bin_data(df, x_range=XR, y_range=YR, grid_res=None, val_col="z")
    with pytest.raises(AssertionError):
        DataPrep.bin_data(df.iloc[:0], x_range=XR, y_range=YR, grid_res=RES, val_col="z")


@pytest.mark.gpu
def test_gpu_bin_data_by_matches_oracle():
    from gpsat_b200.dataprepper import DataPrep
    rng = np.random.default_rng(9)
    df = _frame()
    df["date"] = rng.integers(18320, 18326, len(df)).astype(float)
    df["source"] = rng.choice(["CS2", "S3A", "S3B"], len(df))
    df = df[~((df["date"] == 18322) & (df["source"] == "S3B"))]          # one combination absent from the data
    ref = ob.bin_data_by(df, ["source", "date"], "z", "x", "y", XR, YR, RES)
    out = DataPrep.bin_data_by(df, by_cols=["source", "date"], val_col="z", x_range=XR, y_range=YR, grid_res=RES)
    assert out.index.names == ref.index.names and out.index.equals(ref.index)
    assert np.array_equal(np.isnan(out["z"].values), np.isnan(ref["z"].values))
    np.testing.assert_allclose(out["z"].values, ref["z"].values, rtol=1e-12)
    # several statistics in one call, as examples/bin_data.py:165 asks for (one first pass, one second pass for std)
    multi = DataPrep.bin_data_by(df, by_cols=["source", "date"], val_col="z", x_range=XR, y_range=YR, grid_res=RES,
                                 bin_statistic=["mean", "std", "count", "max"])
    assert list(multi.columns) == ["z_mean", "z_std", "z_count", "z_max"]
    for st, tol in (("mean", 1e-12), ("std", 1e-10), ("count", 0.0), ("max", 0.0)):
        r = ob.bin_data_by(df, ["source", "date"], "z", "x", "y", XR, YR, RES, bin_statistic=st)["z"].values
        o = multi[f"z_{st}"].values
        assert np.array_equal(np.isnan(o), np.isnan(r)), st
        np.testing.assert_allclose(o, r, rtol=tol, atol=1e-15 if tol else 0.0, err_msg=st)
    # the usual follow-up (examples: .dropna().reset_index()) gives the observation table of the hot path
    obs = out.dropna().reset_index()
    assert set(obs.columns) == {"y", "x", "source", "date", "z"} and len(obs) > 1000


@pytest.mark.gpu
def test_gpu_binning_full_size_properties():
    """2e7 raw points (a month of along-track data) on the 5 km pan-Arctic grid: total count = rows inside the
    range, sum of bin sums = sum of the values inside (linearity), mean within [min, max] of the data."""
    from gpsat_b200 import dataprepper as dp
    rng = np.random.default_rng(10)
    n = 20_000_000
    x = rng.uniform(-4.6e6, 4.6e6, n)
    y = rng.uniform(-4.6e6, 4.6e6, n)
    z = rng.normal(0.3, 0.1, n)
    xe, ye = dp._edges([-4.5e6, 4.5e6], [-4.5e6, 4.5e6], 5_000.0, True)
    s, c, _ = dp._accumulate(x, y, z, None, 1, xe, ye)
    inside = (x >= xe[0]) & (x <= xe[-1]) & (y >= ye[0]) & (y <= ye[-1])
    assert int(c.sum()) == int(inside.sum())
    np.testing.assert_allclose(s.sum(), z[inside].sum(), rtol=1e-11)
    m = s[c > 0] / c[c > 0]
    assert m.min() >= z.min() and m.max() <= z.max()
