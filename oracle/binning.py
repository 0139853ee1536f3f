"""TEST INFRASTRUCTURE ONLY: CPU restatement of the upstream binning.

  bin_data     GPSat/dataprepper.py:230-407 -- scipy.stats.binned_statistic_2d on np.linspace edges; pinned against the
               reference's own DataPrep.bin_data outputs (tests/golden/binning.npz, make_golden_binning.py)
  bin_data_by  GPSat/dataprepper.py:23-228 -- one bin_data per unique by_cols combination; the reference assembles an
               xarray Dataset (xarray absent here: frame shaping restated as Dataset.to_dataframe() lays it out, unpinned)
"""
import numpy as np
import pandas as pd
import scipy.stats as scst


def bin_data(df, x_range, y_range, grid_res, x_col="x", y_col="y", val_col=None, bin_statistic="mean", bin_2d=True):
    x_min, x_max = x_range
    y_min, y_max = y_range
    n_x, n_y = int(((x_max - x_min) / grid_res) + 1), int(((y_max - y_min) / grid_res) + 1)     # :340-343
    x_edge, y_edge = np.linspace(x_min, x_max, n_x), np.linspace(y_min, y_max, n_y)                # :352-353
    if bin_2d:
        b = scst.binned_statistic_2d(df[x_col].values, df[y_col].values, df[val_col].values, statistic=bin_statistic,
                                     bins=[x_edge, y_edge], range=[[x_min, x_max], [y_min, y_max]])
    else:
        b = scst.binned_statistic(df[x_col].values, df[val_col].values, statistic=bin_statistic, bins=x_edge,
                                  range=[x_min, x_max])
    xc, yc = x_edge[:-1] + np.diff(x_edge) / 2, y_edge[:-1] + np.diff(y_edge) / 2
    return (b[0].T, (xc, yc)) if bin_2d else (b[0].T, xc)


def bin_data_by(df, by_cols, val_col, x_col, y_col, x_range, y_range, grid_res, bin_statistic="mean"):
    by_cols = [by_cols] if isinstance(by_cols, str) else list(by_cols)
    uniq = [np.sort(df[bc].unique()) for bc in by_cols]
    pieces = {}
    xc = yc = None
    for key, sub in df.groupby(by_cols, sort=True):
        key = key if isinstance(key, tuple) else (key,)
        b, (xc, yc) = bin_data(sub, x_range, y_range, grid_res, x_col, y_col, val_col, bin_statistic)
        pieces[key] = b
    idx = pd.MultiIndex.from_product([yc, xc] + uniq, names=[y_col, x_col] + by_cols)
    full = np.full([len(yc), len(xc)] + [len(u) for u in uniq], np.nan)
    for key, b in pieces.items():
        pos = tuple(int(np.searchsorted(u, k)) for u, k in zip(uniq, key))
        full[(slice(None), slice(None)) + pos] = b
    return pd.DataFrame({val_col: full.reshape(-1)}, index=idx)
