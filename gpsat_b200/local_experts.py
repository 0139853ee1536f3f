"""LocalExpertOI with a batched dispatch: the reference's orchestrator surface over the CUDA engine.

Mirrors ``GPSat.local_experts.LocalExpertOI`` (local_experts.py:116-1463) for the path BASELINE.json names: the
constructor takes the same four config dicts (read through ``gpsat_b200.dataloader``, which restates the slice of
``DataLoader.load`` the reference uses, so configs/example_local_expert_oi.json is accepted unchanged), ``run`` takes
the same arguments and appends the same tables to the same HDF5 file with the same ``HDFStore.append`` keyword
arguments: ``oi_config`` (utils.py:1136-1254), ``expert_locs`` (local_experts.py:882-903), ``run_details``,
``preds`` and one table per hyper-parameter (local_experts.py:499-550), each indexed by the expert's coords_col
values (local_experts.py:691-747), rows in the order the sequential loop (local_experts.py:930-1260) appends them.

What is different by design: consecutive experts are processed in chunks (``max_batch``); inside a chunk all
experts that share a global ``where`` (local_experts.py:426-472) are selected, optimised and predicted in ONE call of
the batched engine instead of one Python model per expert, and the chunk's tables are flushed to the store before the
next chunk starts (the reference flushes every ``store_every`` experts; here a chunk is the flush unit, so a crash
loses at most one chunk and a restart resumes from ``run_details`` exactly like the reference's,
local_experts.py:474-497).  ``load_params={"previous": True}`` (order-dependent EMA warm start,
local_experts.py:1079-1083,1200-1217) and ``replacement_threshold`` (per-expert model switch, :1021-1041) cannot be
batched and are rejected.
"""
from __future__ import annotations

import datetime
import importlib
import json
import os
import re
import sys
import time
import warnings
from ast import literal_eval

import numpy as np
import pandas as pd

from . import dataloader as dl
from .batched import ModelSpec
from .distributed import run_experts_sharded
from .model import B200GPRModel, B200SGPRModel, get_model as _get_model
from .params import HyperParams, PARAM_NAMES


def pretty_print_class(x):
    """GPSat/utils.py:1965-2000"""
    out = str(x)
    if re.search("^<class", out):
        return re.sub("^<class '|'>$", "", out)
    if re.search("^<__main", out):
        return re.sub(r"^<__main__\.| object at .*$", "", out)
    return out


def json_serializable(d, max_len_df=100):
    """dict -> dict that json.dumps accepts (GPSat/utils.py:1331-1434): tuple keys and unknown objects become
    strings, arrays lists, short frames dicts."""
    assert isinstance(d, dict), f"input is type: {type(d)}, expect dict"
    out = {}
    for k, v in d.items():
        if isinstance(k, tuple):
            k = str(k)
        if isinstance(v, dict):
            out[k] = json_serializable(v, max_len_df)
        elif isinstance(v, np.ndarray):
            out[k] = v.tolist()
        elif isinstance(v, (pd.DataFrame, pd.Series)):
            out[k] = json_serializable(v.to_dict(), max_len_df) if len(v) <= max_len_df else str(v)
        else:
            try:
                json.dumps({k: v})
                out[k] = v
            except (TypeError, OverflowError):
                out[k] = str(v)
    return out


def nested_dict_literal_eval(d):
    """inverse of json_serializable's tuple-key stringification (GPSat/utils.py:1276-1328); in place"""
    for k in list(d.keys()):
        if isinstance(k, str) and re.search(r"^\(.*\)$", k):
            try:
                ke = literal_eval(k)
                if ke != k:
                    d[ke] = d.pop(k)
                    k = ke
            except ValueError:
                pass
        if isinstance(d[k], dict):
            d[k] = nested_dict_literal_eval(d[k])
    return d


# ---------------------------------------------------------------------------------------------
# oi_config bookkeeping (GPSat/utils.py:1136-1327)
# ---------------------------------------------------------------------------------------------
def get_previous_oi_config(store_path, oi_config, table_name="oi_config", skip_valid_checks_on=None):
    """Returns (previous config(s), skip_valid_checks_on, config_id) and appends the current config to the
    ``oi_config`` table when it is new (first run: idx 1; no exact match among the stored ones: max idx + 1)."""
    skip_valid_checks_on = [] if skip_valid_checks_on is None else skip_valid_checks_on
    row = pd.DataFrame({"idx": 1, "datetime": datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
                        "config": json.dumps(json_serializable(oi_config))}, index=[1])
    append_kw = dict(index=False, data_columns=["idx", "datetime"], min_itemsize={"config": 50000})
    exists = False
    if os.path.exists(store_path):
        with pd.HDFStore(store_path, mode="r") as store:
            exists = table_name in store
    if not exists:
        with pd.HDFStore(store_path, mode="a") as store:
            store.append(key=table_name, value=row, **append_kw)
            try:
                store.get_storer(table_name).attrs["oi_config"] = oi_config
            except Exception as e:     # the attribute is a convenience copy; the table row is authoritative
                print(e)
        return oi_config, skip_valid_checks_on, 1
    with pd.HDFStore(store_path, mode="a") as store:
        prev = {r["idx"]: nested_dict_literal_eval(json.loads(r["config"])) for _, r in store.get(table_name).iterrows()}
        current = nested_dict_literal_eval(json.loads(row["config"].iloc[0]))
        same = [k for k, v in prev.items() if v == oi_config or v == current]
        if same:
            return prev[max(same)], skip_valid_checks_on, int(max(same))
        row["idx"] = max(prev) + 1
        store.append(key=table_name, value=row, **append_kw)
        return prev, skip_valid_checks_on, int(row["idx"].values[0])


def check_prev_oi_config(prev_oi_config, oi_config, skip_valid_checks_on=None):
    """AssertionError when a key of the current config differs from the previous run's (utils.py:1276-1327 documents
    this; as written the reference's final assert is inverted and a non-matching config dies on a KeyError instead --
    the outcome, "a store is not silently continued with another config", is the same)."""
    skip = [] if skip_valid_checks_on is None else skip_valid_checks_on
    if prev_oi_config == oi_config:
        return
    if prev_oi_config and all(isinstance(k, (int, np.integer)) for k in prev_oi_config):
        prev_oi_config = prev_oi_config[max(prev_oi_config)]        # {idx: config}: compare with the latest
    cur = nested_dict_literal_eval(json.loads(json.dumps(json_serializable(oi_config))))
    bad = [k for k, v in cur.items() if k not in skip and v != prev_oi_config.get(k)]
    assert len(bad) == 0, f"the following keys did not have values that matched exactly: {bad}"


def remove_previously_run_locations(store_path, xprt_locs, table="run_details"):
    """Anti-join of the expert locations with the index of ``table`` in the store (local_experts.py:474-497)."""
    try:
        with pd.HDFStore(store_path, mode="r") as store:
            prev = store.select(table)
        names = list(prev.index.names)
        prev = prev.reset_index()[names].drop_duplicates()
        tmp = xprt_locs.merge(prev, how="left", on=names, indicator="found_already")
        keep = (tmp["found_already"] == "left_only").values
        print(f"for table: {table} returning {keep.sum()} / {len(keep)} entries")
        return xprt_locs.loc[keep].copy(True)
    except (OSError, KeyError) as e:
        print(e)
        return xprt_locs


class LocalExpertOI:
    def __init__(self, expert_loc_config=None, data_config=None, model_config=None, pred_loc_config=None,
                 local_expert_config=None, device: int = 0):
        if local_expert_config is not None and expert_loc_config is None:
            expert_loc_config = local_expert_config
        self.config = {}
        self.device = device
        self.expert_locs = None
        self.model = None
        self.data_source = None
        self.set_expert_locations(**(expert_loc_config or {}))
        self.set_data(**(data_config or {}))
        self.set_model(**(model_config or {}))
        self.set_pred_loc(**(pred_loc_config or {}))

    # ---- config capture (local_experts.py:230-422) ----
    def set_expert_locations(self, df=None, file=None, source=None, where=None, add_data_to_col=None,
                             col_funcs=None, keep_cols=None, col_select=None, row_select=None, sort_by=None,
                             reset_index=False, source_kwargs=None, verbose=False, **kwargs):
        if col_select is None and keep_cols is not None:
            warnings.warn("\n'keep_cols' provided to set_expert_locations, use 'col_select' instead")
            col_select = keep_cols
        if source is None and df is not None:
            warnings.warn("\n'df' was provided to set_expert_locations, use 'source' instead")
            source = df
        if source is None and file is not None:
            warnings.warn("\n'file' was provided to set_expert_locations, use 'source' instead")
            source = file
        if source is None:
            return
        self.config["locations"] = json_serializable(dict(
            df=df, file=file, source=source, where=where, add_data_to_col=add_data_to_col, col_funcs=col_funcs,
            keep_cols=keep_cols, col_select=col_select, row_select=row_select, sort_by=sort_by,
            reset_index=reset_index, source_kwargs=source_kwargs, verbose=verbose, **kwargs))
        locs = dl.load(source=source, where=where, source_kwargs=source_kwargs, col_funcs=col_funcs,
                       row_select=row_select, col_select=col_select, reset_index=reset_index,
                       add_data_to_col=add_data_to_col, **kwargs)
        if sort_by:
            locs = locs.sort_values(sort_by)
        self.expert_locs = locs

    def set_data(self, data_source=None, table=None, obs_col=None, coords_col=None, local_select=None,
                 global_select=None, where=None, row_select=None, col_select=None, col_funcs=None, engine=None,
                 read_kwargs=None, **kwargs):
        self.config["data"] = json_serializable({k: v for k, v in dict(
            data_source=data_source, table=table, obs_col=obs_col, coords_col=coords_col, local_select=local_select,
            global_select=global_select, where=where, row_select=row_select, col_select=col_select,
            col_funcs=col_funcs, engine=engine, read_kwargs=read_kwargs, **kwargs).items()
            if not (v is None and k in ("where", "engine", "read_kwargs"))})
        self.data_table = table
        self.obs_col = obs_col[0] if isinstance(obs_col, (list, tuple)) and len(obs_col) == 1 else obs_col
        self.coords_col = [coords_col] if isinstance(coords_col, str) else coords_col
        self.local_select, self.global_select = local_select, global_select or []
        self.data_where = where
        self.row_select, self.col_select, self.col_funcs = row_select, col_select, col_funcs
        # a path is opened once (HDFStore stays open read-only for the run, like LocalExpertData.set_data_source)
        self.data_source = dl.open_source(data_source, engine, **(read_kwargs or {})) \
            if isinstance(data_source, str) else data_source

    def set_model(self, oi_model=None, init_params=None, constraints=None, load_params=None, optim_kwargs=None,
                  pred_kwargs=None, params_to_store=None, replacement_threshold=None, replacement_model=None,
                  replacement_init_params=None, replacement_constraints=None, replacement_optim_kwargs=None,
                  replacement_pred_kwargs=None):
        self.config["model"] = json_serializable(dict(
            oi_model=oi_model, init_params=init_params, constraints=constraints, load_params=load_params,
            optim_kwargs=optim_kwargs, pred_kwargs=pred_kwargs, params_to_store=params_to_store,
            replacement_threshold=replacement_threshold, replacement_model=replacement_model,
            replacement_init_params=replacement_init_params, replacement_constraints=replacement_constraints,
            replacement_optim_kwargs=replacement_optim_kwargs, replacement_pred_kwargs=replacement_pred_kwargs))
        if oi_model is None:
            self.model = None
            return
        if isinstance(oi_model, str):
            self.model = _get_model(oi_model)
        elif isinstance(oi_model, dict):         # the reference's custom-model hook (local_experts.py:319-325)
            path = oi_model["path_to_model"]
            sys.path.append(path)
            self.model = getattr(importlib.import_module(path), oi_model["model_name"])
        else:
            self.model = oi_model
        assert self.model in (B200GPRModel, B200SGPRModel), "the batched driver dispatches B200GPRModel / B200SGPRModel"

        def mc(model, ip, cons, ok):
            return dict(init_params=ip or {}, constraints=cons, optim_kwargs=ok or {},
                        oi_model="B200SGPRModel" if model is B200SGPRModel else "B200GPRModel")

        self.model_config = mc(self.model, init_params, constraints, optim_kwargs)
        self.pred_kwargs = pred_kwargs or {}
        assert not self.pred_kwargs.get("full_cov", False), "full_cov predictions are not stored by run()"
        self.load_params_config = load_params
        self.params_to_store = None if params_to_store == "all" else params_to_store
        # Two features make experts order-dependent or pick the model per expert -- load_params={"previous": True}
        # (EMA warm start, local_experts.py:1079-1083,1200-1217) and replacement_threshold (:1021-1041).  They are
        # served by the sequential fallback of run(): the same engine, one expert per call, in list order.
        self.replacement_threshold = replacement_threshold
        self.replacement_model_config = None
        if replacement_threshold is not None:
            rm = self.model if replacement_model is None else _get_model(replacement_model)
            assert rm in (B200GPRModel, B200SGPRModel), "replacement_model must be a gpsat_b200 model"
            self.replacement_model = rm
            self.replacement_model_config = mc(rm, init_params if replacement_init_params is None else
                                               replacement_init_params,
                                               constraints if replacement_constraints is None else
                                               replacement_constraints, replacement_optim_kwargs)
            assert not (replacement_pred_kwargs or {}).get("full_cov", False)

    def set_pred_loc(self, method="expert_loc", coords_col=None, df=None, df_file=None, max_dist=None,
                     copy_df=False, **kwargs):
        self.config["pred_loc"] = json_serializable(dict(method=method, coords_col=coords_col, df=df, df_file=df_file,
                                                         max_dist=max_dist, copy_df=copy_df, **kwargs))
        assert method in ("expert_loc", "from_dataframe"), f"pred_loc method '{method}' is not handled"
        self.pred_method, self.pred_max_dist = method, max_dist
        self.pred_df = None
        if method == "from_dataframe":
            if df is None:       # prediction_locations.py:213-218
                assert isinstance(df_file, str), f"df is None, df_file expected to be str, got: {type(df_file)}"
                df = pd.read_csv(df_file)
            self.pred_df = df

    # ---- parameter loading (local_experts.py:553-689), vectorised over experts ----
    def _same_param_table(self, store_path, table_suffix):
        """local_experts.py:749-759: loading from and (not optimising) writing to the same table"""
        lp = self.load_params_config
        extra = [k for k in lp if k not in ("file", "table_suffix")]
        return (store_path == lp.get("file", None)) and (table_suffix == lp.get("table_suffix", None)) \
            and len(extra) == 0

    def _load_theta(self, locs: pd.DataFrame):
        """Per-expert start values [E, D+2] and a mask of experts that can run.

        ``load_params`` forms (local_experts.py:553-606): ``file`` (+ ``table_suffix``, ``param_names``,
        ``index_adjust``) -> one lookup per parameter table keyed on the expert's coords_col values; otherwise the
        remaining keys are fixed values applied to every expert (``set_parameters(**param_dict)``).  A parameter that
        is missing or NaN for an expert keeps the model default; an expert for which the file yields nothing at all
        is skipped (status 1, local_experts.py:595-596,1099-1101)."""
        lp = self.load_params_config
        E, D = len(locs), len(self.coords_col)
        if lp is None:
            return None, np.ones(E, dtype=bool)
        spec = ModelSpec.from_model_config(self.model_config)
        default = HyperParams(D, spec.lengthscales, spec.kernel_variance, spec.likelihood_variance).theta()
        theta = np.tile(default, (E, 1))
        sl = {"lengthscales": slice(0, D), "kernel_variance": slice(D, D + 1),
              "likelihood_variance": slice(D + 1, D + 2)}
        src = lp.get("file", None)
        if src is None:
            fixed = {k: v for k, v in lp.items() if k not in ("previous", "previous_params", "file", "param_names",
                                                              "ref_loc", "index_adjust", "table_suffix")}
            for nm, v in fixed.items():
                assert nm in sl, f"cannot set parameter '{nm}': not one of {list(sl)}"
                theta[:, sl[nm]] = np.broadcast_to(np.asarray(v, dtype=np.float64).ravel(), (sl[nm].stop - sl[nm].start,))
            return theta, np.ones(E, dtype=bool)
        suffix = lp.get("table_suffix", "")
        names = lp.get("param_names") or list(PARAM_NAMES)
        for nm in names:
            assert nm in sl, f"provide param name:{nm}\nis not in param_names:{list(sl)} handled by the batched loader"
        key = locs[self.coords_col].reset_index(drop=True).copy()
        for c, spec_c in (lp.get("index_adjust") or {}).items():        # e.g. parameters of the previous day
            key[c] = [dl.config_func(**spec_c, args=v) for v in key[c].values]
        key["_row_"] = np.arange(E)
        found_any = np.zeros(E, dtype=bool)
        if isinstance(src, str):
            assert os.path.exists(src), f"in load_params file provided:\n{src}\nbut path does not exist"
        for nm in names:
            try:
                if isinstance(src, dict):           # in-memory tables (tests, chained runs)
                    tab = src[f"{nm}{suffix}"]
                else:
                    with pd.HDFStore(src, mode="r") as store:
                        tab = store.select(f"{nm}{suffix}")
            except KeyError as e:
                print("KeyError\n", e, f"\nskipping param_name: {nm}")
                continue
            tab = tab.reset_index()
            if "_dim_0" not in tab.columns:
                tab["_dim_0"] = 0
            width = sl[nm].stop - sl[nm].start
            m = key.merge(tab[self.coords_col + ["_dim_0", nm]], how="inner", on=self.coords_col)
            m = m[(m["_dim_0"] >= 0) & (m["_dim_0"] < width)]
            vals = np.full((E, width), np.nan)
            vals[m["_row_"].values, m["_dim_0"].values.astype(int)] = m[nm].values
            got = ~np.isnan(vals).any(axis=1)         # a NaN anywhere drops the parameter for that expert
            theta[got, sl[nm]] = vals[got]
            found_any |= got
        return theta, found_any

    # ---- the run ----
    def run(self, store_path=None, store_every=10, check_config_compatible=True, skip_valid_checks_on=None,
            optimise=True, predict=True, min_obs=3, table_suffix="", return_tables=None, max_batch=4096):
        """Same arguments as the reference's ``run`` (local_experts.py:761-769) plus ``max_batch`` (experts per
        engine chunk = flush unit) and ``return_tables``.  Tables are appended to the HDF5 file at ``store_path``
        (pandas.HDFStore, needs PyTables like the reference); with ``store_path=None`` or ``return_tables=True``
        they are returned as a dict of DataFrames."""
        from . import get_engine
        import torch.distributed as dist
        t_run = time.perf_counter()
        assert isinstance(self.expert_locs, pd.DataFrame), \
            f"attr expert_locs is {type(self.expert_locs)}, expected to be DataFrame"
        assert self.data_source is not None, "'data_source' is None"
        assert self.model is not None, "'model' is None"
        min_obs, store_every, max_batch = int(min_obs), int(store_every), int(max_batch)
        assert store_every >= 1, f"store_every must be >= 1, got: {store_every}"
        assert min_obs >= 1, f"min_obs must be >= 1, got: {min_obs}"
        assert max_batch >= 1
        if return_tables is None:
            return_tables = store_path is None
        self.config["run_kwargs"] = json_serializable(dict(
            store_path=store_path, store_every=store_every, check_config_compatible=check_config_compatible,
            skip_valid_checks_on=skip_valid_checks_on, optimise=optimise, predict=predict, min_obs=min_obs,
            table_suffix=table_suffix))
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        is_writer = (not multi) or dist.get_rank() == 0
        coords_col = self.coords_col
        out_tables = {}

        # ---- store bookkeeping: config id, expert_locs table, resume (local_experts.py:862-912) ----
        config_id, keep_pos = 1, None
        if store_path is not None and is_writer:
            os.makedirs(os.path.dirname(os.path.abspath(store_path)), exist_ok=True)
            prev_cfg, skip_valid_checks_on, config_id = get_previous_oi_config(
                store_path, self.config, skip_valid_checks_on=skip_valid_checks_on,
                table_name=f"oi_config{table_suffix}")
            if check_config_compatible:
                check_prev_oi_config(prev_cfg, self.config, skip_valid_checks_on)
            store_locs = remove_previously_run_locations(store_path, self.expert_locs.copy(True),
                                                         table=f"expert_locs{table_suffix}")
            store_locs = store_locs.set_index(coords_col)
            with pd.HDFStore(store_path, mode="a") as store:
                store.append(f"expert_locs{table_suffix}", store_locs, data_columns=True)
            left = remove_previously_run_locations(store_path, self.expert_locs.copy(True).assign(
                _pos_=np.arange(len(self.expert_locs))), table=f"run_details{table_suffix}")
            keep_pos = left["_pos_"].values
        if multi:      # one reader / writer of the store: rank 0 decides, everybody follows
            box = [config_id, keep_pos]
            dist.broadcast_object_list(box, src=0)
            config_id, keep_pos = box
        xprt = self.expert_locs.copy(True)
        if keep_pos is not None:
            xprt = xprt.iloc[keep_pos]
        if store_path is None:
            out_tables[f"expert_locs{table_suffix}"] = self.expert_locs.set_index(coords_col)
            out_tables[f"oi_config{table_suffix}"] = pd.DataFrame(
                {"idx": config_id, "datetime": datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
                 "config": json.dumps(json_serializable(self.config))}, index=[config_id])

        eng = get_engine(self.device)
        spec = ModelSpec.from_model_config(self.model_config)
        D = len(coords_col)
        model_name = pretty_print_class(self.model)[:64]
        dev_name = self._device_name()[:64]
        sel_cols = [c for ls in self.local_select for c in ([ls["col"]] if isinstance(ls["col"], str) else ls["col"])]
        table_cols = list(dict.fromkeys(list(coords_col) + [self.obs_col] + sel_cols))
        pred_cols, pred_tab = None, None
        if self.pred_method == "from_dataframe":
            pred_cols = [c for c in coords_col if c in self.pred_df.columns]
            pred_tab = np.ascontiguousarray(self.pred_df[pred_cols].values.T, dtype=np.float64)
        # only the columns the kernels read travel to the device (the C ABI takes at most 16 reference columns)
        ref_cols = [c for c in dict.fromkeys(list(coords_col) + sel_cols)
                    if c in xprt.columns and np.issubdtype(xprt[c].dtype, np.number)]
        assert all(c in ref_cols for c in coords_col), "expert locations must hold every (numeric) coords_col"
        save_params = not (self.load_params_config is not None and (not optimise)
                           and self._same_param_table(store_path, table_suffix))
        rows = xprt.to_dict("records")
        cache_key, cache_df = None, None
        previous = bool((self.load_params_config or {}).get("previous", False))
        if previous or self.replacement_threshold is not None:
            assert not multi, "the sequential fallback (previous / replacement_threshold) runs on one GPU"
            self._run_sequential(eng, xprt, rows, table_cols, ref_cols, pred_tab, pred_cols, store_path, store_every,
                                 optimise, predict, min_obs, table_suffix, model_name, dev_name, config_id, D,
                                 save_params, previous, out_tables if return_tables else None)
            xprt = xprt.iloc[:0]      # nothing left for the batched loop
        for c0 in range(0, len(xprt), max_batch):
            members_all = np.arange(c0, min(c0 + max_batch, len(xprt)))
            # group the chunk's experts by their global where list (local_experts.py:426-472)
            groups, order = {}, []
            for i in members_all:
                w = dl.get_where_list(self.global_select, self.local_select, rows[i])
                k = json.dumps(w, default=str, sort_keys=True)
                if k not in groups:
                    groups[k] = (w, [])
                    order.append(k)
                groups[k][1].append(i)
            pieces = {}
            for k in order:
                where, members = groups[k]
                if k != cache_key:
                    base = ([self.data_where] if isinstance(self.data_where, dict) else list(self.data_where or []))
                    cache_df = dl.load(source=self.data_source, table=self.data_table, where=base + where,
                                       col_funcs=self.col_funcs, row_select=self.row_select,
                                       col_select=self.col_select, reset_index=True)
                    cache_key = k
                gdf = cache_df
                t0 = time.perf_counter()
                if len(gdf):
                    table = np.ascontiguousarray(gdf[table_cols].values.T, dtype=np.float64)
                else:
                    table = np.zeros((len(table_cols), 1)) + np.inf     # nothing can be selected
                sub = xprt.iloc[members]
                theta_init, ok_load = self._load_theta(sub)
                kw = dict(pred_table=pred_tab, pred_cols=pred_cols, max_dist=self.pred_max_dist, min_obs=min_obs)
                res, res_bad = None, None
                if ok_load.any():
                    refs = np.ascontiguousarray(sub[ref_cols].values[ok_load], dtype=np.float64)
                    # one process per GPU: the expert list is sharded by N^3 cost and gathered once (distributed.py)
                    res = run_experts_sharded(eng, spec, table, table_cols, self.obs_col, coords_col, refs, ref_cols,
                                              self.local_select, optimise=optimise, predict=predict,
                                              theta_init=None if theta_init is None else theta_init[ok_load], **kw)
                if not ok_load.all():
                    # experts whose parameters could not be loaded are not run (local_experts.py:1099-1101), but the
                    # ones with too few observations are still recorded (the min_obs test comes first, :988-1012)
                    refs = np.ascontiguousarray(sub[ref_cols].values[~ok_load], dtype=np.float64)
                    res_bad = run_experts_sharded(eng, spec, table, table_cols, self.obs_col, coords_col, refs,
                                                  ref_cols, self.local_select, count_only=True, **kw)
                dt = time.perf_counter() - t0
                self._shape_tables(pieces, res, res_bad, sub, np.asarray(members), ok_load, dt, optimise, predict,
                                   model_name, dev_name, config_id, D, save_params)
            chunk = {}
            # tables in the order the reference's save_dict holds them (run_details, preds, parameters)
            for name, lst in sorted(pieces.items(), key=lambda kv: {"run_details": 0, "preds": 1}.get(kv[0], 2)):
                df = lst[0] if len(lst) == 1 else pd.concat(lst, axis=0)
                # rows back in the order the sequential loop would have appended them (one where-group per chunk, the
                # usual case, is in that order already: no second pass over the prediction rows)
                posv = df["_pos_"].values
                if len(posv) > 1 and (posv[1:] < posv[:-1]).any():
                    df = df.iloc[np.argsort(posv, kind="stable")]
                chunk[f"{name}{table_suffix}"] = df.drop(columns="_pos_")
            if store_path is not None and is_writer:
                self._flush(store_path, chunk)
            if return_tables:
                for name, df in chunk.items():
                    out_tables.setdefault(name, []).append(df)
        print(f"'run': {time.perf_counter() - t_run:.3f} seconds")
        if not return_tables:
            return None
        return {k: (pd.concat(v, axis=0) if isinstance(v, list) else v) for k, v in out_tables.items()}

    def _global_frame(self, where, cache):
        """the observations passing the global where list, cached while consecutive experts share it"""
        k = json.dumps(where, default=str, sort_keys=True)
        if cache.get("key") != k:
            base = ([self.data_where] if isinstance(self.data_where, dict) else list(self.data_where or []))
            cache["df"] = dl.load(source=self.data_source, table=self.data_table, where=base + where,
                                  col_funcs=self.col_funcs, row_select=self.row_select, col_select=self.col_select,
                                  reset_index=True)
            cache["key"] = k
        return cache["df"]

    def _run_sequential(self, eng, xprt, rows, table_cols, ref_cols, pred_tab, pred_cols, store_path, store_every,
                        optimise, predict, min_obs, table_suffix, model_name, dev_name, config_id, D, save_params,
                        previous, out_tables):
        """The reference's loop order for the two order-dependent features (local_experts.py:930-1260): one expert per
        engine call; ``previous`` starts every expert from the exponential moving average (rho = 0.95) of the
        parameters of the experts that optimised successfully before it (:1200-1217); ``replacement_threshold``
        switches to the replacement model's settings when an expert has fewer observations (:1021-1041)."""
        spec_main = ModelSpec.from_model_config(self.model_config)
        spec_repl = None if self.replacement_model_config is None else \
            ModelSpec.from_model_config(self.replacement_model_config)
        prev = None                     # EMA of theta [D+2]
        pending, n_pending = {}, 0
        cache = {}
        is_writer = True

        def flush():
            nonlocal pending, n_pending
            chunk = {}
            for name, lst in sorted(pending.items(), key=lambda kv: {"run_details": 0, "preds": 1}.get(kv[0], 2)):
                chunk[f"{name}{table_suffix}"] = pd.concat(lst, axis=0).drop(columns="_pos_")
            if store_path is not None and is_writer:
                self._flush(store_path, chunk)
            if out_tables is not None:
                for name, df in chunk.items():
                    out_tables.setdefault(name, []).append(df)
            pending, n_pending = {}, 0

        for i in range(len(xprt)):
            sub = xprt.iloc[[i]]
            where = dl.get_where_list(self.global_select, self.local_select, rows[i])
            gdf = self._global_frame(where, cache)
            table = np.ascontiguousarray(gdf[table_cols].values.T, dtype=np.float64) if len(gdf) else \
                np.zeros((len(table_cols), 1)) + np.inf
            refs = np.ascontiguousarray(sub[ref_cols].values, dtype=np.float64)
            kw = dict(pred_table=pred_tab, pred_cols=pred_cols, max_dist=self.pred_max_dist, min_obs=min_obs)
            spec, mname = spec_main, model_name
            if spec_repl is not None:
                cnt = run_experts_sharded(eng, spec_main, table, table_cols, self.obs_col, self.coords_col, refs,
                                          ref_cols, self.local_select, count_only=True, **kw)
                if int(cnt["num_obs"][0]) < self.replacement_threshold:
                    spec, mname = spec_repl, pretty_print_class(self.replacement_model)[:64]
            theta_init, ok = None, np.ones(1, dtype=bool)
            if previous:
                if prev is None:        # the first model's parameters as constructed (local_experts.py:1051-1052)
                    prev = HyperParams(D, spec.lengthscales, spec.kernel_variance, spec.likelihood_variance).theta()
                theta_init = prev[None, :].copy()
            elif self.load_params_config is not None:
                theta_init, ok = self._load_theta(sub)
            t0 = time.perf_counter()
            pieces = {}
            if ok[0]:
                res = run_experts_sharded(eng, spec, table, table_cols, self.obs_col, self.coords_col, refs, ref_cols,
                                          self.local_select, optimise=optimise, predict=predict,
                                          theta_init=theta_init, **kw)
                res_bad = None
            else:
                res = None
                res_bad = run_experts_sharded(eng, spec, table, table_cols, self.obs_col, self.coords_col, refs,
                                              ref_cols, self.local_select, count_only=True, **kw)
            self._shape_tables(pieces, res, res_bad, sub, np.array([i]), ok, time.perf_counter() - t0, optimise,
                               predict, mname, dev_name, config_id, D, save_params)
            if res is not None and res.get("n_valid", 0) and optimise and int(res["status"][0]) in (1, 2):
                th = np.asarray(res["theta"][0], dtype=np.float64)
                if prev is not None and not np.isnan(th).any():
                    prev = 0.95 * prev + 0.05 * th
            if pieces:
                for name, lst in pieces.items():
                    pending.setdefault(name, []).extend(lst)
                n_pending += 1
                if n_pending >= store_every:
                    flush()
        if n_pending:
            flush()

    @classmethod
    def run_from(cls, ref_oi, **run_kwargs):
        """Batched run of an already configured reference ``GPSat.local_experts.LocalExpertOI`` instance
        (the hook shown in INTEGRATION.md): its captured config dicts rebuild the driver."""
        cfg = ref_oi.config
        model_cfg = dict(cfg.get("model") or {})
        oi = cls(expert_loc_config=None, data_config=cfg.get("data"), model_config=model_cfg,
                 pred_loc_config=cfg.get("pred_loc"))
        locs_cfg = cfg.get("locations") or cfg.get("local_expert_locations")
        if getattr(ref_oi, "expert_locs", None) is not None:
            oi.expert_locs = ref_oi.expert_locs
            oi.config["locations"] = locs_cfg
        elif locs_cfg:
            oi.set_expert_locations(**locs_cfg)
        data = getattr(ref_oi, "data", None)
        if data is not None and isinstance(getattr(data, "data_source", None), pd.DataFrame):
            oi.data_source = data.data_source         # in-memory sources are not part of the JSON-able config
        pl = getattr(ref_oi, "pred_loc", None)
        if pl is not None and isinstance(getattr(pl, "kwargs", {}).get("df", None), pd.DataFrame):
            oi.pred_df = pl.kwargs["df"]
        return oi.run(**run_kwargs)

    def _device_name(self):
        import torch
        return torch.cuda.get_device_name(self.device)

    def _shape_tables(self, pieces, res, res_bad, sub, pos, ok_load, dt, optimise, predict, model_name, dev_name,
                      config_id, D, save_params):
        """Vectorised dict_of_array_to_table (local_experts.py:691-747) for a whole batch."""
        coords_col = self.coords_col
        ref_all = sub[coords_col].values

        def midx(rows_ref, repeats=None):
            """index of the expert coordinates, row e repeated repeats[e] times.  The levels are factorised over the
            EXPERT rows and the codes repeated (MultiIndex.from_arrays over the repeated rows gives the same index but
            factorises every level over millions of prediction rows: a third of the host time of a predict-only batch)"""
            if len(coords_col) == 1:
                v = rows_ref[:, 0] if repeats is None else np.repeat(rows_ref[:, 0], repeats)
                return pd.Index(v, name=coords_col[0])
            levels, codes = [], []
            for j in range(len(coords_col)):
                cat = pd.Categorical(rows_ref[:, j])               # sorted unique levels, as from_arrays builds them
                levels.append(cat.categories)
                c = np.asarray(cat.codes)
                codes.append(c if repeats is None else np.repeat(c, repeats))
            return pd.MultiIndex(levels=levels, codes=codes, names=coords_col, verify_integrity=False)

        def details(r, num_obs, rt, fobj, succ, ran, ref, pp):
            return pd.DataFrame({"_dim_0": 0, "num_obs": num_obs[r], "run_time": rt[r],
                                 "objective_value": fobj[r], "parameters_optimised": optimise,
                                 "optimise_success": succ[r], "model": model_name,
                                 "device": np.where(ran[r], dev_name, ""), "config_id": config_id,
                                 "_pos_": pp[r]}, index=midx(ref[r]))

        if res_bad is not None:         # parameters missing: only the "too few observations" rows are recorded
            b = np.flatnonzero(~ok_load)
            r = np.flatnonzero(res_bad["too_few"])
            if len(r):
                nanv, f = np.full(len(b), np.nan), np.zeros(len(b), dtype=bool)
                pieces.setdefault("run_details", []).append(
                    details(r, res_bad["num_obs"], nanv, nanv, f, f, ref_all[b], pos[b]))
        if res is None:
            return
        g = np.flatnonzero(ok_load)
        ref, pp = ref_all[g], pos[g]
        E = len(g)
        num_obs, too_few = res["num_obs"], res["too_few"]
        valid_idx = res.get("valid_idx", np.zeros(0, dtype=np.int64))
        Ev = len(valid_idx)
        keep = np.zeros(E, dtype=bool)
        keep[valid_idx] = True
        vpos = np.full(E, -1)
        vpos[valid_idx] = np.arange(Ev)
        fobj = np.full(E, np.nan)
        succ = np.zeros(E, dtype=bool)
        if Ev:
            fobj[valid_idx] = res["fobj"]
            if optimise:
                succ[valid_idx] = np.isin(res["status"], (1, 2))
        rt = np.where(keep, dt / max(int(keep.sum()), 1), np.nan)
        r = np.flatnonzero(keep | too_few)
        if len(r):
            pieces.setdefault("run_details", []).append(details(r, num_obs, rt, fobj, succ, keep, ref, pp))
        k = np.flatnonzero(keep)
        if len(k) == 0:
            return
        # hyper-parameter tables (skipped when loading from and writing to the same table without optimising)
        if save_params:
            th = res["theta"][vpos[k]]
            names = self.params_to_store or (list(PARAM_NAMES) + ["inducing_points"])
            if "lengthscales" in names:
                pieces.setdefault("lengthscales", []).append(pd.DataFrame(
                    {"_dim_0": np.tile(np.arange(D), len(k)), "lengthscales": th[:, :D].ravel(),
                     "_pos_": np.repeat(pp[k], D)}, index=midx(ref[k], D)))
            if "kernel_variance" in names:
                pieces.setdefault("kernel_variance", []).append(pd.DataFrame(
                    {"_dim_0": 0, "kernel_variance": th[:, D], "_pos_": pp[k]}, index=midx(ref[k])))
            if "likelihood_variance" in names:
                pieces.setdefault("likelihood_variance", []).append(pd.DataFrame(
                    {"_dim_0": 0, "likelihood_variance": th[:, D + 1], "_pos_": pp[k]}, index=midx(ref[k])))
            if "inducing_points" in res and "inducing_points" in names:
                zo, zc = res["z_offsets"], res["inducing_points"]
                mk = np.diff(zo)[vpos[k]]
                rows = np.concatenate([np.arange(zo[v], zo[v + 1]) for v in vpos[k]])
                d0 = np.concatenate([np.repeat(np.arange(m), D) for m in mk])
                pieces.setdefault("inducing_points", []).append(pd.DataFrame(
                    {"_dim_0": d0, "_dim_1": np.tile(np.arange(D), len(rows)), "inducing_points": zc[rows].ravel(),
                     "_pos_": np.repeat(pp[k], mk * D)}, index=midx(ref[k], mk * D)))
        if predict:
            # The prediction table is millions of rows for a predict-only batch: its float columns are gathered straight
            # into ONE (columns x rows) block that the frame adopts without copying (a dict of column arrays is stacked
            # and consolidated by pandas: two more passes over ~100 MB), and the rows of consecutive experts are a slice.
            poff = res["pred_offsets"]
            cntk = np.diff(poff)[vpos[k]]
            starts = np.asarray(poff)[vpos[k]]
            n_rows = int(cntk.sum())
            first_row = np.cumsum(cntk) - cntk
            dim0 = np.arange(n_rows) - np.repeat(first_row, cntk)
            if np.array_equal(starts[1:], starts[:-1] + cntk[:-1]):
                sl = slice(int(starts[0]), int(starts[0]) + n_rows)

                def take(a, out):
                    np.copyto(out, a[sl])
            else:
                sel = dim0 + np.repeat(starts, cntk)

                def take(a, out):
                    np.take(a, sel, out=out)
            local_mean = self.model_config["init_params"].get("obs_mean", None) == "local"
            fcols = ["f*", "f*_var", "y_var"] + (["f_bar"] if local_mean else []) + [f"pred_loc_{c}" for c in coords_col]
            blk = np.empty((len(fcols), n_rows), dtype=np.float64)
            take(res["fmean"], blk[0])
            take(res["fvar"], blk[1])
            take(res["yvar"], blk[2])
            if local_mean:
                blk[3] = np.repeat(res["obs_mean"][vpos[k]], cntk)
            for ci in range(len(coords_col)):
                take(res["pred_coords"][:, ci], blk[len(fcols) - len(coords_col) + ci])
            pr = pd.DataFrame(blk.T, columns=fcols, index=midx(ref[k], cntk), copy=False)
            pr.insert(0, "_dim_0", dim0)
            if not local_mean:
                # base_model.py:199-200: without obs_mean='local' the mean is the INTEGER array [[0]], and predict
                # broadcasts it (gpflow_models.py:265-271): the stored column is int64 zeros, not float
                pr.insert(4, "f_bar", np.zeros(n_rows, dtype=np.int64))
            pr["_pos_"] = np.repeat(pp[k], cntk)
            pieces.setdefault("preds", []).append(pr)

    @staticmethod
    def _flush(store_path, tables):
        """One HDFStore.append per table with the reference's keyword arguments (local_experts.py:535-543)."""
        for k, v in tables.items():
            min_itemsize = {c: 64 for c in v.columns if c in ["model", "device"]}
            try:
                with pd.HDFStore(store_path, mode="a") as store:     # raises ImportError without PyTables
                    store.append(key=k, value=v, min_itemsize=min_itemsize)
            except ValueError as e:
                print(e)


def get_results_from_h5file(results_file, global_col_funcs=None, merge_on_expert_locations=True,
                            select_tables=None, table_suffix="", add_suffix_to_table=True, verbose=False):
    """(dict of result tables with the index reset, list of oi_config dicts) from a results file
    (local_experts.py:1467-1620); expert-location columns are merged onto every table that holds the coords_col."""
    if select_tables is not None and add_suffix_to_table:
        select_tables = [f"{t}{table_suffix}" for t in select_tables]
    with pd.HDFStore(results_file, mode="r") as store:
        try:
            cfg = store.get(f"oi_config{table_suffix}")[["config"]].drop_duplicates()
            oi_config = [nested_dict_literal_eval(json.loads(c)) for c in cfg["config"].values]
        except Exception:
            try:
                oi_config = [nested_dict_literal_eval(store.get_storer(f"oi_config{table_suffix}").attrs["oi_config"])]
            except Exception:
                oi_config = []
        keys = [re.sub("^/", "", k) for k in store.keys()]
        wanted = keys if select_tables is None else select_tables
        dfs = {}
        for k in keys:
            if k in wanted:
                try:
                    dfs[k] = store.select(k).reset_index()
                except Exception as e:
                    print(f"issue with key: {k}\n{e}")
    if global_col_funcs is not None:
        for k in dfs:
            try:
                dl.add_cols(dfs[k], global_col_funcs)
            except Exception as e:
                print(f"Adding/Modifying columns had Exception:{e}\non key/table: {k}")
    expert_locations = None
    if f"expert_locs{table_suffix}" in dfs:
        expert_locations = dfs[f"expert_locs{table_suffix}"].copy(True)
    else:
        try:
            expert_locations = pd.concat([LocalExpertOI(expert_loc_config=c["locations"]).expert_locs.copy(True)
                                          for c in oi_config])
        except Exception as e:
            print(f"in get_results_from_h5file trying read expert_locations from file got Exception:\n{e}")
    if expert_locations is not None and merge_on_expert_locations:
        try:
            coords_col = oi_config[0]["data"]["coords_col"]
        except KeyError:
            coords_col = oi_config[0]["input_data"]["coords_col"]
        for k in dfs:
            if np.isin(coords_col, dfs[k].columns).all():
                dfs[k] = dfs[k].merge(expert_locations, on=coords_col, how="left", suffixes=["", "_expert_location"])
    return dfs, oi_config


def get_results_from_tables(tables, table_suffix=""):
    """Tables returned by ``run(return_tables=True)`` with the suffix stripped from their names."""
    return {re.sub(f"{re.escape(table_suffix)}$", "", k): v for k, v in tables.items()}
