"""Post-processing either side of the hot path, on the GPU.

Mirrors of the reference functions (same names, argument meaning, output frames):

  gaussian_2d_weight          GPSat/postprocessing.py:22-52   (numba gufunc)
  smooth_hyperparameter_table the per-parameter loop of smooth_hyperparameters, postprocessing.py:240-308
                              (every (t, _dim_*) slice in ONE kernel launch)
  get_weighted_values         GPSat/utils.py:2081-2214
  glue_local_predictions_1d / _2d   GPSat/postprocessing.py:447-577

The sums run in hand-written CUDA (csrc/postproc.cuh through gpsat_gaussian_smooth / gpsat_weighted_groups);
grouping / sorting of the keys and the DataFrame shaping stay on the host.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import List, Union

import numpy as np
import pandas as pd
import torch

from . import _lib


def _dev(a, dtype, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device)


def _device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("gpsat_b200.postprocessing needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", device)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _smooth_segments(qx, qy, seg_of_q, x, y, vals, seg_off, l_x, l_y, vmin=None, vmax=None, device=0):
    lib = _lib.load()
    dev = _device(device)
    nq = len(qx)
    out = torch.empty(nq, dtype=torch.float64, device=dev)
    if nq == 0:
        return out.cpu().numpy()
    t = [_dev(qx, np.float64, dev), _dev(qy, np.float64, dev), _dev(seg_of_q, np.int32, dev),
         _dev(x, np.float64, dev), _dev(y, np.float64, dev), _dev(vals, np.float64, dev),
         _dev(seg_off, np.int64, dev)]
    cmin = None if vmin is None else C.byref(C.c_double(float(vmin)))
    cmax = None if vmax is None else C.byref(C.c_double(float(vmax)))
    _lib.check(lib.gpsat_gaussian_smooth(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), nq, t[3].data_ptr(),
                                         t[4].data_ptr(), t[5].data_ptr(), t[6].data_ptr(), float(l_x), float(l_y),
                                         cmin, cmax, out.data_ptr(), _stream(dev)))
    return out.cpu().numpy()


def gaussian_2d_weight(x0, y0, x, y, l_x, l_y, vals, device=0):
    """out[i] = sum_j w_ij vals_j / sum_j w_ij, w_ij = exp(-(((x_j-x0_i)/l_x)^2 + ((y_j-y0_i)/l_y)^2)/2),
    NaN vals skipped, NaN when no weight is left (postprocessing.py:22-52; core dims `(),(),(n),(n),(),(),(n)->()`
    with x0 / y0 broadcast over the leading axis, as smooth_hyperparameters calls it)."""
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
    y0 = np.atleast_1d(np.asarray(y0, dtype=np.float64))
    x, y, vals = (np.asarray(a, dtype=np.float64).ravel() for a in (x, y, vals))
    assert x0.shape == y0.shape and x.shape == y.shape == vals.shape
    seg_off = np.array([0, len(x)], dtype=np.int64)
    return _smooth_segments(x0, y0, np.zeros(len(x0), np.int32), x, y, vals, seg_off, l_x, l_y, device=device)


def smooth_hyperparameter_table(df: pd.DataFrame, val_col: str, coords_col: List[str], l_x, l_y, max=None, min=None,
                                xy_dims=("x", "y"), device=0) -> pd.DataFrame:
    """One parameter table (columns: coords, optional _dim_*, `val_col`) -> its smoothed table, indexed by
    `coords_col`, exactly as the loop body of smooth_hyperparameters builds it (postprocessing.py:240-308):
    slices = unique combinations of the non-xy coordinates and `_dim_*` columns in order of first appearance,
    values clipped to [min, max] before weighting, rows whose smoothed value is NaN dropped."""
    xy_dims = list(xy_dims)
    x_col, y_col = xy_dims
    org_cols = df.columns.values.tolist()
    other_dims = [c for c in coords_col if c not in xy_dims] + [c for c in df.columns if re.search(r"^_dim_\d", c)]
    if len(other_dims):
        key = df[other_dims].apply(tuple, axis=1) if len(other_dims) > 1 else df[other_dims[0]]
        seg = pd.factorize(key, sort=False)[0].astype(np.int64)       # order of first appearance (drop_duplicates)
    else:
        seg = np.zeros(len(df), dtype=np.int64)
    order = np.argsort(seg, kind="stable")                              # merge(how="inner") keeps df's row order
    seg_s = seg[order]
    G = int(seg_s[-1]) + 1 if len(seg_s) else 0
    seg_off = np.zeros(G + 1, dtype=np.int64)
    np.cumsum(np.bincount(seg_s, minlength=G), out=seg_off[1:])
    x = df[x_col].values[order].astype(np.float64)
    y = df[y_col].values[order].astype(np.float64)
    vals = df[val_col].values[order].astype(np.float64)
    sm = _smooth_segments(x, y, seg_s.astype(np.int32), x, y, vals, seg_off, l_x, l_y, vmin=min, vmax=max,
                          device=device)
    out = df.iloc[order].copy()
    out[val_col] = sm
    out = out[~np.isnan(sm)]
    out = out[org_cols]
    return out.set_index(list(coords_col))


def _weighted_groups(ref, to, vals, lengthscale, device=0):
    """ref, to: (n, nd); vals: (n, ncol).  Returns (unique ref rows sorted lexicographically (G, nd),
    weighted means (G, ncol), sum of weights (G,))."""
    lib = _lib.load()
    dev = _device(device)
    ref = np.asarray(ref, dtype=np.float64).reshape(len(ref), -1)
    to = np.asarray(to, dtype=np.float64).reshape(len(to), -1)
    vals = np.asarray(vals, dtype=np.float64).reshape(len(ref), -1)
    n, nd = ref.shape
    ncol = vals.shape[1]
    order = np.lexsort(ref.T[::-1]).astype(np.int64)                    # groupby / pivot_table sort by the keys
    rs = ref[order]
    new = np.ones(n, dtype=bool)
    if n:
        new[1:] = np.any(rs[1:] != rs[:-1], axis=1)
    starts = np.flatnonzero(new)
    G = len(starts)
    off = np.concatenate([starts, [n]]).astype(np.int64)
    out = torch.empty((ncol + 1, G), dtype=torch.float64, device=dev)
    if G:
        t = [_dev(ref.T, np.float64, dev), _dev(to.T, np.float64, dev), _dev(vals.T, np.float64, dev),
             _dev(order, np.int64, dev), _dev(off, np.int64, dev)]
        _lib.check(lib.gpsat_weighted_groups(t[0].data_ptr(), t[1].data_ptr(), nd, t[2].data_ptr() if ncol else None,
                                             n, ncol, t[3].data_ptr(), t[4].data_ptr(), G, float(lengthscale),
                                             out.data_ptr(), _stream(dev)))
    o = out.cpu().numpy()
    return rs[starts], o[:ncol].T, o[ncol]


def get_weighted_values(df, ref_col, dist_to_col, val_cols, weight_function="gaussian", drop_weight_cols=True,
                        device=0, **weight_kwargs):
    """GPSat/utils.py:2081-2214: per unique `ref_col` row, sum(w v)/sum(w) of each `val_cols` column with
    w = exp(-(|ref - dist_to|^2 / lengthscale^2) / 2); output sorted by the reference columns."""
    ref_col = [ref_col] if isinstance(ref_col, str) else list(ref_col)
    dist_to_col = [dist_to_col] if isinstance(dist_to_col, str) else list(dist_to_col)
    val_cols = [val_cols] if isinstance(val_cols, str) else list(val_cols)
    x0 = df[ref_col].values
    x = df[dist_to_col].values
    assert x0.shape == x.shape, \
        f"ref_col gave shape: {x0.shape}, dist_to_col gave shape: {x.shape} - they should be the same"
    if weight_function != "gaussian":
        raise NotImplementedError(f"weight_function: {weight_function} is not implemented")
    lscale = weight_kwargs.get("lengthscale", None)
    assert lscale is not None, "lscale is None, please provide"
    keys, means, wsum = _weighted_groups(x0, x, df[val_cols].values, lscale, device=device)
    out = pd.DataFrame({c: keys[:, i].astype(df[c].dtype, copy=False) for i, c in enumerate(ref_col)})
    for i, vc in enumerate(val_cols):
        if not drop_weight_cols:
            out["_w"] = wsum
            out[f"w_{vc}"] = means[:, i] * wsum
        out[vc] = means[:, i]
    return out


def _glue(preds_df, pred_loc_cols, xprt_loc_cols, vars_to_glue, inference_radius, R, device):
    if isinstance(vars_to_glue, str):
        vars_to_glue = [vars_to_glue]
    assert isinstance(inference_radius, (int, float)), "a per-expert inference_radius dict is not supported here"
    # prod_k norm.pdf(p_k, e_k, s) = exp(-|p - e|^2 / (2 s^2)) / (2 pi s^2)^(d/2): the constant cancels in the ratio
    keys, means, _ = _weighted_groups(preds_df[pred_loc_cols].values, preds_df[xprt_loc_cols].values,
                                      preds_df[vars_to_glue].values, inference_radius / R, device=device)
    out = pd.DataFrame({c: keys[:, i] for i, c in enumerate(pred_loc_cols)})
    for i, v in enumerate(vars_to_glue):
        out[v] = means[:, i]
    return out


def glue_local_predictions_1d(preds_df: pd.DataFrame, pred_loc_col: str, xprt_loc_col: str,
                              vars_to_glue: Union[str, List[str]], inference_radius, R=3, device=0) -> pd.DataFrame:
    """GPSat/postprocessing.py:447-515 (scalar inference_radius)."""
    return _glue(preds_df, [pred_loc_col], [xprt_loc_col], vars_to_glue, inference_radius, R, device)


def glue_local_predictions_2d(preds_df: pd.DataFrame, pred_loc_cols: List[str], xprt_loc_cols: List[str],
                              vars_to_glue: Union[str, List[str]], inference_radius, R=3, device=0) -> pd.DataFrame:
    """GPSat/postprocessing.py:518-577."""
    return _glue(preds_df, list(pred_loc_cols), list(xprt_loc_cols), vars_to_glue, inference_radius, R, device)
