"""Selection oracle (numpy).  TEST INFRASTRUCTURE ONLY.

Restates, per SURVEY.md section 8a:
  S1  DataLoader.get_where_list / _bool_numpy_from_where  (GPSat/dataloader.py:2892-2978,1886-1971)
  S2  DataLoader.local_data_select                        (GPSat/dataloader.py:2352-2447)
  S3  PredictionLocations._from_dataframe/_max_dist_bool  (GPSat/prediction_locations.py:18-43,208-273)

S2's multi-column branch is scipy.spatial.KDTree.query_ball_point (p=2, eps=0): a point is
returned iff  sum_j (x_j - c_j)^2 <= r*r  evaluated in float64, squares summed in column
order with no FMA contraction -- inclusive although the configs spell the comparison "<".
S3 is strict: every per-dimension (d*d) < max_dist^2 AND the row sum < max_dist^2.
"""
from __future__ import annotations

import operator

import numpy as np
import pandas as pd

_COMP = {">=": operator.ge, ">": operator.gt, "==": operator.eq, "<": operator.lt, "<=": operator.le,
         "!=": operator.ne}


def where_mask(cols: dict, where_list) -> np.ndarray:
    """AND of static where-dicts {"col","comp","val"} over a dict of equal-length columns (S1)."""
    n = len(next(iter(cols.values())))
    m = np.ones(n, dtype=bool)
    for w in where_list or []:
        m &= _COMP[w["comp"]](cols[w["col"]], w["val"])
    return m


def expand_where_list(global_select, local_select, ref_loc: dict):
    """dataloader.py:2892-2978: static dicts pass through, dynamic dicts are expanded per local_select."""
    out = []
    for gs in global_select or []:
        if all(c in gs for c in ("col", "comp", "val")):
            out.append(gs)
            continue
        assert all(c in gs for c in ("loc_col", "src_col", "func"))
        func = gs["func"]
        if isinstance(func, str):
            func = eval(func, {"np": np, "pd": pd})  # noqa: S307 - same contract as the reference's config lambdas
        for ls in local_select:
            if gs["loc_col"] == ls["col"]:
                out.append({"col": gs["src_col"], "comp": ls["comp"],
                            "val": func(ref_loc[gs["loc_col"]], ls["val"])})
    return out


def local_select_mask(cols: dict, ref_loc: dict, local_select) -> np.ndarray:
    """Boolean mask over rows, AND of the local_select entries in listed order (S2)."""
    n = len(next(iter(cols.values())))
    select = np.ones(n, dtype=bool)
    for ls in local_select:
        col, comp = ls["col"], ls["comp"]
        if isinstance(col, str):
            assert comp in (">=", ">", "==", "<", "<=")
            select &= _COMP[comp](cols[col], ref_loc[col] + ls["val"])
        else:
            assert comp in ("<", "<=")
            d2 = np.zeros(n)
            for c in col:  # column order, plain mul+add (no FMA in numpy)
                d = cols[c] - ref_loc[c]
                d2 = d2 + d * d
            select &= d2 <= ls["val"] * ls["val"]
    return select


def local_select_indices(cols, ref_loc, local_select):
    return np.flatnonzero(local_select_mask(cols, ref_loc, local_select))


def max_dist_mask(locs: np.ndarray, ref: np.ndarray, max_dist: float) -> np.ndarray:
    """prediction_locations.py:18-43: strict per-dimension prefilter then strict squared-L2."""
    md2 = max_dist * max_dist
    out = np.ones(len(locs), dtype=bool)
    for j in range(locs.shape[1]):
        d = locs[:, j] - ref[j]
        out &= (d * d) < md2
    d2 = np.zeros(len(locs))
    for j in range(locs.shape[1]):
        d = locs[:, j] - ref[j]
        d2 = d2 + d * d
    out &= d2 < md2
    return out


def prediction_locations(pred_cols: dict, coords_col, expert_row: dict, max_dist=None):
    """_from_dataframe: (P, D) array; coords missing from the frame take the expert's value."""
    found = [c for c in coords_col if c in pred_cols]
    fc_loc = [coords_col.index(c) for c in found]
    locs = np.stack([np.asarray(pred_cols[c], dtype=np.float64) for c in found], axis=1)
    ref = np.array([expert_row[c] for c in coords_col], dtype=np.float64)
    if max_dist is not None:
        b = max_dist_mask(locs, ref[fc_loc], max_dist)
        locs = locs[b]
    else:
        b = np.ones(len(locs), dtype=bool)
    out = np.full((len(locs), len(coords_col)), np.nan)
    out[:, fc_loc] = locs
    for j, c in enumerate(coords_col):
        if c not in found:
            out[:, j] = ref[j]
    return out, np.flatnonzero(b)
