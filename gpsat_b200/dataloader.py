"""Host-side loading of expert locations / observations for the batched LocalExpertOI driver.

Mirrors the slice of ``GPSat.dataloader.DataLoader`` that ``LocalExpertOI`` goes through, so that a reference
config (e.g. configs/example_local_expert_oi.json) is read unchanged:

  load             DataLoader.load                       GPSat/dataloader.py:1522-1680
  modify_df        DataLoader._modify_df                 dataloader.py:1682-1799   (order: add_data_to_col -> col_funcs ->
                                                                                    row_select -> col_select)
  add_data_to_col  DataLoader.add_data_to_col            dataloader.py:1416-1499   (a list value cross-joins the frame)
  add_cols         DataLoader.add_cols + config_func     dataloader.py:46-134, GPSat/utils.py:311-493
  where_mask       DataLoader._bool_numpy_from_where     dataloader.py:1886-1971
  get_where_list   DataLoader.get_where_list             dataloader.py:2892-2978
  store_where      DataLoader._hdfstore_where_from_dict  dataloader.py:1839-1851   (global select pushed into HDFStore.select)

Everything here is pandas on the host: it runs once per run / per global-where group, not per expert.
"""
from __future__ import annotations

import importlib
import os
import re
from functools import reduce

import numpy as np
import pandas as pd

_SUFFIX_ENGINE = {"csv": "read_csv", "tsv": "read_csv", "h5": "HDFStore", "parquet": "read_parquet",
                  "pq": "read_parquet", "pkl": "read_pickle", "pickle": "read_pickle"}
_COMPS = (">=", ">", "==", "<", "<=")


# ---------------------------------------------------------------------------------------------
# functions named in configs
# ---------------------------------------------------------------------------------------------
def _column(df, col, as_numpy=True):
    try:
        s = df.loc[:, col]
    except KeyError:
        assert isinstance(col, int), f"col: {col} not a column name, and isn't an integer"
        s = df.iloc[:, col]
    if not as_numpy:
        return s
    v = s.values
    # pandas >= 3 keeps strings in an extension array; config lambdas such as x.astype('datetime64[D]') are written
    # for the numpy object arrays older pandas returned
    return v if isinstance(v, np.ndarray) else s.to_numpy(dtype=object)


def config_func(func, source=None, args=None, kwargs=None, col_args=None, col_kwargs=None, df=None,
                filename_as_arg=False, filename=None, col_numpy=True):
    """Apply a function described in a JSON config (GPSat/utils.py:311-493): ``func`` is a callable, a
    ``"lambda ..."`` string, a binary operator string (``"<="``), or a name importable from ``source``;
    ``col_args`` / ``col_kwargs`` name columns of ``df`` passed (as numpy arrays) ahead of ``args`` / ``kwargs``."""
    args = [] if args is None else (args if isinstance(args, list) else [args])
    col_args = [] if col_args is None else (col_args if isinstance(col_args, list) else [col_args])
    kwargs = {} if kwargs is None else kwargs
    col_kwargs = {} if col_kwargs is None else col_kwargs
    assert isinstance(kwargs, dict), "kwargs needs to be a dict"
    assert isinstance(col_kwargs, dict), "col_kwargs needs to be a dict"
    if df is None:
        assert len(col_args) == 0, f"df not provide, but col_args: {col_args} were"
        assert len(col_kwargs) == 0, f"df not provide, but col_kwargs: {col_kwargs} were"
    else:
        col_args = [_column(df, c, col_numpy) for c in col_args]
        col_kwargs = {k: _column(df, c, col_numpy) for k, c in col_kwargs.items()}
    call_args = col_args + args
    if filename_as_arg and filename is not None:
        call_args = [filename] + call_args
    call_kwargs = {**col_kwargs, **kwargs}
    if isinstance(func, str):
        env = {"np": np, "pd": pd}
        if re.search("^lambda", func):
            fun = eval(func, env)                                   # noqa: S307  the reference's config contract
        elif re.search(r"[\|&\=\+\-\*/\%<>]", func):
            fun = lambda a, b: eval(f"a {func} b", env, {"a": a, "b": b})   # noqa: S307,E731
        else:
            try:
                fun = eval(func, env)                               # noqa: S307
            except NameError:
                assert source is not None, f"NameError occurred on eval({func}), cannot import"
                fun = getattr(importlib.import_module(source), func)
    else:
        assert callable(func), "func provided is not str nor is it callable"
        fun = func
    out = fun(*call_args, **call_kwargs)
    return out.values if isinstance(out, pd.Series) else out


# ---------------------------------------------------------------------------------------------
# row / column manipulation
# ---------------------------------------------------------------------------------------------
def add_data_to_col(df, add=None):
    """Scalar -> constant column; list -> the frame is repeated once per entry (dataloader.py:1416-1499)."""
    add = {} if add is None else add
    assert isinstance(add, dict), f"add_cols expected to be dict, got: {type(add)}"
    for col, vals in add.items():
        if isinstance(vals, (int, str, float)):
            vals = [vals]
        parts = []
        for v in vals:
            part = df.copy(True)
            part[col] = v
            parts.append(part)
        df = pd.concat(parts, axis=0)
    return df


def add_cols(df, col_funcs=None, filename=None):
    """In place: new column <- config_func(df=df, **spec); a tuple key takes one returned array per name."""
    for new_col, spec in (col_funcs or {}).items():
        vals = config_func(df=df, filename=filename, **spec)
        if isinstance(new_col, tuple):
            assert len(vals) == len(new_col), \
                f"columns: {list(new_col)} have length: {len(new_col)} but function returned {len(vals)} values"
            for name, v in zip(new_col, vals):
                df[name] = v
        else:
            df[new_col] = vals


def where_mask(df, wd):
    """One condition: {"col","comp","val"} is evaluated on the pandas Series (so a string compares against a
    datetime column the way pandas does it); anything else is a config_func spec; "negate" flips the result."""
    wd = dict(wd)
    negate = wd.pop("negate", False)
    if all(k in wd for k in ("col", "comp", "val")):
        col, comp, val = wd["col"], wd["comp"], wd["val"]
        assert isinstance(df, (pd.Series, pd.DataFrame))
        assert col in df.columns, f"col: '{col}' is not in coords: {df.columns}"
        assert comp in _COMPS, f"comp: {comp} is not valid"
        out = eval(f"x {comp} y", {}, {"x": df[col], "y": val})     # noqa: S307
    else:
        out = config_func(df=df, **wd)
    return ~out if negate else out


def row_select_bool(df, row_select=None, combine="AND"):
    if row_select is None:
        row_select = []
    elif isinstance(row_select, dict):
        row_select = [row_select]
    assert isinstance(row_select, list), f"expect row_select to be a list (of dict), is type: {type(row_select)}"
    for i, rs in enumerate(row_select):
        assert isinstance(rs, dict), f"index element: {i} of row_select was type: {type(rs)}, rather than dict"
    combine = combine.upper()
    assert combine in ("AND", "OR"), f"combine: {combine} not in ['AND','OR']"
    masks = [where_mask(df, wd) for wd in row_select]
    if not masks:
        return slice(None)
    return reduce((lambda a, b: a & b) if combine == "AND" else (lambda a, b: a | b), masks)


def modify_df(df, col_funcs=None, filename=None, row_select=None, col_select=None, add_data_to_col_=None,
              combine_row_select="AND"):
    df = add_data_to_col(df, add_data_to_col_)
    add_cols(df, col_funcs, filename=filename)
    df = df.loc[row_select_bool(df, row_select, combine_row_select), :]
    if col_select is None:
        return df
    missing = [c for c in col_select if c not in df]
    assert len(missing) == 0, f"columns were provide, but {missing} are not in obj (dataframe)"
    return df.loc[:, col_select]


# ---------------------------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------------------------
def store_where(wd):
    """{"col","comp","val"} -> the where-string handed to HDFStore.select (dataloader.py:1839-1851)."""
    val = wd["val"]
    if isinstance(val, str) or isinstance(val, (np.datetime64, pd.Timestamp)):
        val = f'"{val}"'
    elif isinstance(val, (int, float, bool, list)):
        val = str(val)
    return "".join([wd["col"], wd["comp"], val])


def open_source(source, engine=None, **kwargs):
    """A path becomes a reader: HDFStore for .h5 (needs PyTables, like the reference), pandas.read_* otherwise."""
    if not isinstance(source, str):
        return source
    if engine is None:
        suffix = re.sub(r"^.*\.", "", source)
        assert suffix in _SUFFIX_ENGINE, f"file_suffix: {suffix} not in file_suffix_engine_map: {_SUFFIX_ENGINE}"
        engine = _SUFFIX_ENGINE[suffix]
    if engine == "HDFStore":
        return pd.HDFStore(source, mode="r", **kwargs)
    assert hasattr(pd, engine) and engine.startswith("read"), f"engine: {engine} was not understood"
    return getattr(pd, engine)(source, **kwargs)


def _as_list(where):
    if isinstance(where, dict):
        where = [where]
    return where or None


def data_select(obj, where=None, table=None, columns=None, reset_index=False, combine_where="AND", close=False):
    """Rows of ``obj`` (DataFrame / dict / open HDFStore) matching the list of where-dicts (dataloader.py:1011-1273)."""
    where = _as_list(where)
    combine_where = combine_where.upper()
    assert combine_where in ("AND", "OR"), f"combine_where='{combine_where}' is not valid, must be either 'AND' or 'OR'"
    if isinstance(obj, pd.io.pytables.HDFStore):
        assert table is not None, "\n\nobj is HDFStore, however table is None, needs to be provided\n\n"
        w = [store_where(wd) for wd in where] if where else None
        try:
            if combine_where == "AND" or w is None:
                out = obj.select(key=table, where=w, columns=columns)
            else:
                out = pd.concat([obj.select(key=table, where=x, columns=columns) for x in w], axis=0)
        finally:
            if close:
                obj.close()
        if reset_index:
            out.reset_index(inplace=True)
        return out
    if isinstance(obj, dict):
        obj = pd.DataFrame(obj)
    assert isinstance(obj, (pd.DataFrame, pd.Series)), f"type(obj): {type(obj)} was not understood"
    rows = slice(None)
    if where:
        masks = [where_mask(obj, wd) for wd in where]
        rows = reduce((lambda a, b: a & b) if combine_where == "AND" else (lambda a, b: a | b), masks)
    if columns is not None:
        missing = [c for c in columns if c not in obj]
        assert len(missing) == 0, f"columns were provide, but {missing} are not in obj (dataframe)"
    return obj.loc[rows, slice(None) if columns is None else columns].copy()


def load(source, where=None, engine=None, table=None, source_kwargs=None, col_funcs=None, row_select=None,
         col_select=None, reset_index=False, add_data_to_col=None, close=False, verbose=False,
         combine_row_select="AND", **kwargs):
    """DataLoader.load: read ``source`` (frame, open store or path), apply ``where`` at the source, then add columns
    and select rows / columns in memory."""
    if isinstance(source, str):
        close = True
        source = open_source(source, engine, **(source_kwargs or {}))
    df = data_select(source, where=where, table=table, reset_index=reset_index, close=close, **kwargs)
    return modify_df(df, col_funcs=col_funcs, row_select=row_select, col_select=col_select,
                     add_data_to_col_=add_data_to_col, combine_row_select=combine_row_select)


# ---------------------------------------------------------------------------------------------
# global select
# ---------------------------------------------------------------------------------------------
def get_where_list(global_select, local_select=None, ref_loc=None):
    """Static {"col","comp","val"} entries pass through; dynamic {"loc_col","src_col","func"} entries expand once
    per local_select entry on ``loc_col`` into {"col": src_col, "comp": ls.comp, "val": func(ref[loc_col], ls.val)}."""
    out = []
    for gs in global_select or []:
        if all(k in gs for k in ("col", "comp", "val")):
            out.append(gs)
            continue
        assert local_select is not None, f"dynamic where provide: {gs}, however local_select is: {type(local_select)}"
        assert ref_loc is not None, f"dynamic where provide: {gs}, however ref_loc is: {type(ref_loc)}"
        assert all(k in gs for k in ("loc_col", "src_col", "func")), \
            f"dynamic where had keys: {gs.keys()}, must have: ['loc_col', 'src_col', 'func'] "
        loc_col = gs["loc_col"]
        assert loc_col in ref_loc, f"loc_col: {loc_col} not in ref_loc: {ref_loc}"
        func = gs["func"]
        if isinstance(func, str):
            func = eval(func, {"np": np, "pd": pd})                 # noqa: S307
        for ls in local_select:
            if ls["col"] == loc_col:
                out.append({"col": gs["src_col"], "comp": ls["comp"], "val": func(ref_loc[loc_col], ls["val"])})
    return out


def file_exists(path):
    return isinstance(path, str) and os.path.exists(path)
