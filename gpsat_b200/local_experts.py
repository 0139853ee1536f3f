"""LocalExpertOI with a batched dispatch: the reference's orchestrator surface over the CUDA engine.

Mirrors ``GPSat.local_experts.LocalExpertOI`` (local_experts.py:116-1463) for the path BASELINE.json
names: the constructor takes the same four config dicts, ``run`` takes the same arguments and emits
the same tables (``run_details``, ``preds``, one table per hyper-parameter, ``expert_locs``,
``oi_config``), each indexed by the expert's coords_col values (local_experts.py:691-747), in the
same expert order the sequential loop (local_experts.py:930-1260) would append them.

What is different by design: all experts that share a global ``where`` (local_experts.py:426-472)
are selected, optimised and predicted in ONE call of the batched engine instead of one Python model
per expert.  ``load_params={"previous": True}`` (order-dependent EMA warm start,
local_experts.py:1079-1083,1200-1217) cannot be batched and is rejected.
"""
from __future__ import annotations

import datetime
import json
import os
import re
import time

import numpy as np
import pandas as pd

from .batched import ModelSpec
from .distributed import run_experts_sharded
from .model import B200GPRModel, B200SGPRModel, get_model as _get_model
from .params import PARAM_NAMES

_COMP = {">=": np.greater_equal, ">": np.greater, "==": np.equal, "<": np.less, "<=": np.less_equal,
         "!=": np.not_equal}


def pretty_print_class(x):
    """GPSat/utils.py:1965-2000"""
    out = str(x)
    if re.search("^<class", out):
        return re.sub("^<class '|'>$", "", out)
    if re.search("^<__main", out):
        return re.sub(r"^<__main__\.| object at .*$", "", out)
    return out


def _load_frame(source, table=None):
    if isinstance(source, pd.DataFrame):
        return source
    assert isinstance(source, str), f"source must be a DataFrame or a file path, got: {type(source)}"
    ext = os.path.splitext(source)[1].lower()
    if ext == ".csv":
        return pd.read_csv(source)
    if ext in (".parquet", ".pq"):
        return pd.read_parquet(source)
    if ext in (".h5", ".hdf5"):
        return pd.read_hdf(source, key=table)      # needs PyTables, like the reference
    if ext in (".pkl", ".pickle"):
        return pd.read_pickle(source)
    raise NotImplementedError(f"file type of '{source}' is not handled")


def _apply_where(df: pd.DataFrame, where_list):
    """AND of static {"col","comp","val"} dicts (DataLoader._bool_numpy_from_where, dataloader.py:1886-1971)."""
    if not where_list:
        return df
    m = np.ones(len(df), dtype=bool)
    for w in where_list:
        col = df[w["col"]].values
        val = w["val"]
        if np.issubdtype(col.dtype, np.datetime64) and not isinstance(val, np.datetime64):
            val = np.datetime64(val)
        m &= _COMP[w["comp"]](col, val)
    return df.loc[m]


def _where_list(global_select, local_select, ref_row: dict):
    """DataLoader.get_where_list (dataloader.py:2892-2978)."""
    out = []
    for gs in global_select or []:
        is_static = all(c in gs for c in ("col", "comp", "val"))
        is_dynamic = all(c in gs for c in ("loc_col", "src_col", "func"))
        assert is_static or is_dynamic, f"global_select entry not understood: {gs}"
        if is_static:
            out.append(dict(gs))
            continue
        func = gs["func"]
        if isinstance(func, str):
            func = eval(func, {"np": np, "pd": pd})  # noqa: S307  same contract as the reference's config lambdas
        for ls in local_select:
            if gs["loc_col"] == ls["col"]:
                out.append({"col": gs["src_col"], "comp": ls["comp"], "val": func(ref_row[gs["loc_col"]], ls["val"])})
    return out


def _json_default(o):
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, pd.DataFrame):
        return f"<DataFrame shape={o.shape}>"
    return str(o)


class LocalExpertOI:
    def __init__(self, expert_loc_config=None, data_config=None, model_config=None, pred_loc_config=None,
                 local_expert_config=None, device: int = 0):
        if local_expert_config is not None and expert_loc_config is None:
            expert_loc_config = local_expert_config
        self.config = {}
        self.device = device
        self.set_expert_locations(**(expert_loc_config or {}))
        self.set_data(**(data_config or {}))
        self.set_model(**(model_config or {}))
        self.set_pred_loc(**(pred_loc_config or {}))

    # ---- config capture (local_experts.py:230-422) ----
    def set_expert_locations(self, df=None, file=None, source=None, where=None, add_data_to_col=None,
                             col_funcs=None, keep_cols=None, col_select=None, row_select=None, sort_by=None,
                             reset_index=False, source_kwargs=None, verbose=False, **kwargs):
        self.config["locations"] = {k: v for k, v in dict(file=file, source=source, where=where,
                                                          row_select=row_select, sort_by=sort_by).items()
                                    if v is not None and not isinstance(v, pd.DataFrame)}
        src = df if df is not None else (source if source is not None else file)
        if src is None:
            self.expert_locs = None
            return
        locs = _load_frame(src).copy()
        for k, v in (add_data_to_col or {}).items():
            locs[k] = v
        rs = row_select if row_select is not None else where
        if rs:
            locs = _apply_where(locs, rs if isinstance(rs, list) else [rs])
        cs = col_select if col_select is not None else keep_cols
        if cs:
            locs = locs[cs]
        if sort_by:
            locs = locs.sort_values(sort_by)
        if reset_index:
            locs = locs.reset_index(drop=True)
        self.expert_locs = locs

    def set_data(self, data_source=None, table=None, obs_col=None, coords_col=None, local_select=None,
                 global_select=None, row_select=None, col_select=None, col_funcs=None, engine=None,
                 read_kwargs=None, **kwargs):
        assert col_funcs is None, "col_funcs are not applied by the batched driver: pre-compute the columns"
        self.data_source, self.data_table = data_source, table
        self.obs_col = obs_col[0] if isinstance(obs_col, (list, tuple)) and len(obs_col) == 1 else obs_col
        self.coords_col = [coords_col] if isinstance(coords_col, str) else coords_col
        self.local_select, self.global_select = local_select, global_select or []
        self.row_select, self.col_select = row_select, col_select
        self.config["data"] = {k: v for k, v in dict(data_source=data_source, table=table, obs_col=obs_col,
                                                     coords_col=coords_col, local_select=local_select,
                                                     global_select=global_select, row_select=row_select,
                                                     col_select=col_select).items()
                               if not isinstance(v, pd.DataFrame)}

    def set_model(self, oi_model=None, init_params=None, constraints=None, load_params=None, optim_kwargs=None,
                  pred_kwargs=None, params_to_store=None, replacement_threshold=None, replacement_model=None,
                  replacement_init_params=None, replacement_constraints=None, replacement_optim_kwargs=None,
                  replacement_pred_kwargs=None):
        self.config["model"] = dict(oi_model=oi_model, init_params=init_params, constraints=constraints,
                                    load_params=load_params, optim_kwargs=optim_kwargs, pred_kwargs=pred_kwargs,
                                    params_to_store=params_to_store)
        if oi_model is None:
            self.model = None
            return
        if isinstance(oi_model, str):
            self.model = _get_model(oi_model)
        elif isinstance(oi_model, dict):
            import importlib
            self.model = getattr(importlib.import_module(oi_model["path_to_model"]), oi_model["model_name"])
        else:
            self.model = oi_model
        assert self.model in (B200GPRModel, B200SGPRModel), "the batched driver dispatches B200GPRModel / B200SGPRModel"
        assert replacement_threshold is None, "replacement models are not supported by the batched dispatch"
        self.model_config = dict(init_params=init_params or {}, constraints=constraints,
                                 optim_kwargs=optim_kwargs or {},
                                 oi_model="B200SGPRModel" if self.model is B200SGPRModel else "B200GPRModel")
        self.pred_kwargs = pred_kwargs or {}
        assert not self.pred_kwargs.get("full_cov", False), "full_cov predictions are not stored by run()"
        self.load_params_config = load_params
        if load_params is not None:
            assert not load_params.get("previous", False), \
                "load_params={'previous': True} makes experts order-dependent and cannot be batched"
        self.params_to_store = params_to_store

    def set_pred_loc(self, method="expert_loc", coords_col=None, df=None, df_file=None, max_dist=None,
                     copy_df=False, **kwargs):
        self.config["pred_loc"] = {k: v for k, v in dict(method=method, df_file=df_file, max_dist=max_dist).items()
                                   if v is not None}
        assert method in ("expert_loc", "from_dataframe"), f"pred_loc method '{method}' is not handled"
        self.pred_method, self.pred_max_dist = method, max_dist
        self.pred_df = None
        if method == "from_dataframe":
            self.pred_df = df if df is not None else _load_frame(df_file)

    # ---- parameter loading (local_experts.py:553-689), vectorised over experts ----
    def _load_theta(self, locs: pd.DataFrame, store_tables=None):
        lp = self.load_params_config
        if lp is None:
            return None, np.ones(len(locs), dtype=bool)
        D = len(self.coords_col)
        theta = np.full((len(locs), D + 2), np.nan)
        suffix = lp.get("table_suffix", "")
        src = lp.get("file", None)
        tables = store_tables if (src is None and store_tables is not None) else None
        sl = {"lengthscales": slice(0, D), "kernel_variance": slice(D, D + 1),
              "likelihood_variance": slice(D + 1, D + 2)}
        names = lp.get("param_names") or PARAM_NAMES
        key = locs[self.coords_col].reset_index(drop=True)
        key["_row_"] = np.arange(len(key))
        for nm in names:
            if isinstance(src, dict):
                df = src[f"{nm}{suffix}"]
            elif tables is not None:
                df = tables[f"{nm}{suffix}"]
            else:
                df = pd.read_hdf(src, key=f"{nm}{suffix}")
            df = df.reset_index()
            m = key.merge(df, how="left", on=self.coords_col)
            if "_dim_0" in m.columns:
                m = m.sort_values(["_row_", "_dim_0"])
            vals = m[nm].values.reshape(len(key), -1)
            theta[:, sl[nm]] = vals
        ok = ~np.isnan(theta[:, [sl[n].start for n in names]]).any(axis=1)
        # parameters not loaded keep the model defaults
        spec = ModelSpec.from_model_config(self.model_config)
        from .params import HyperParams
        d = HyperParams(D, spec.lengthscales, spec.kernel_variance, spec.likelihood_variance).theta()
        for j in range(D + 2):
            col = theta[:, j]
            col[np.isnan(col) & ok] = d[j]
        return theta, ok

    # ---- the run ----
    def run(self, store_path=None, store_every=10, check_config_compatible=True, skip_valid_checks_on=None,
            optimise=True, predict=True, min_obs=3, table_suffix="", return_tables=None):
        """Same arguments as the reference's ``run`` (local_experts.py:761-769).  Tables are appended to the
        HDF5 file at ``store_path`` (pandas.HDFStore, needs PyTables like the reference); with
        ``store_path=None`` or ``return_tables=True`` they are returned as a dict of DataFrames."""
        from . import get_engine
        t_run = time.perf_counter()
        assert self.model is not None, "'model' is None"
        assert self.expert_locs is not None, "expert locations were not provided"
        min_obs, store_every = int(min_obs), int(store_every)
        assert min_obs >= 1, f"min_obs must be >= 1, got: {min_obs}"
        if return_tables is None:
            return_tables = store_path is None
        eng = get_engine(self.device)
        spec = ModelSpec.from_model_config(self.model_config)
        coords_col, obs_col = self.coords_col, self.obs_col
        D = len(coords_col)
        self.config["run_kwargs"] = dict(optimise=optimise, predict=predict, min_obs=min_obs,
                                         table_suffix=table_suffix)
        config_id = 1
        src = _load_frame(self.data_source, self.data_table)
        if self.row_select:
            src = _apply_where(src, self.row_select)
        xprt = self.expert_locs.copy(True)
        # resume: drop experts already present in run_details (local_experts.py:474-497, 905-912)
        prev_tables = None
        if store_path is not None and os.path.exists(store_path):
            try:
                with pd.HDFStore(store_path, mode="r") as st:
                    if f"/run_details{table_suffix}" in st.keys():
                        prev = st.get(f"run_details{table_suffix}").reset_index()[coords_col]
                        tmp = xprt.merge(prev.drop_duplicates(), how="left", on=coords_col, indicator="found_already")
                        xprt = xprt.loc[(tmp["found_already"] == "left_only").values].copy(True)
                    if f"/oi_config{table_suffix}" in st.keys():
                        config_id = int(st.get(f"oi_config{table_suffix}")["idx"].max()) + 1
            except Exception as e:      # same spirit as the reference's try/except-and-print
                print(e)
        # group experts by their global where list (local_experts.py:426-472)
        rows = xprt.to_dict("records")
        groups, order = {}, []
        for i, r in enumerate(rows):
            w = _where_list(self.global_select, self.local_select, r)
            k = json.dumps(w, default=_json_default, sort_keys=True)
            if k not in groups:
                groups[k] = (w, [])
                order.append(k)
            groups[k][1].append(i)
        model_name = pretty_print_class(self.model)[:64]
        dev_name = self._device_name()[:64]
        table_cols = list(dict.fromkeys(list(coords_col) + [obs_col] + [c for ls in self.local_select
                                                                         for c in ([ls["col"]] if isinstance(ls["col"], str) else ls["col"])]))
        ref_cols = [c for c in xprt.columns if np.issubdtype(xprt[c].dtype, np.number)]
        pred_cols, pred_tab = None, None
        if self.pred_method == "from_dataframe":
            pred_cols = [c for c in coords_col if c in self.pred_df.columns]
            pred_tab = np.ascontiguousarray(self.pred_df[pred_cols].values.T, dtype=np.float64)
        pieces = {}          # table name -> list of (first expert position, DataFrame)
        for k in order:
            where, members = groups[k]
            gdf = _apply_where(src, where).reset_index(drop=True)
            if self.col_select:
                gdf = gdf[self.col_select]
            t0 = time.perf_counter()
            table = np.ascontiguousarray(gdf[table_cols].values.T, dtype=np.float64)
            sub = xprt.iloc[members]
            refs = np.ascontiguousarray(sub[ref_cols].values, dtype=np.float64)
            theta_init, ok_load = self._load_theta(sub)
            if len(gdf) == 0:
                table = np.zeros((len(table_cols), 1)) + np.inf     # nothing can be selected
            # one process per GPU: the expert list is sharded by N^3 cost and gathered once (distributed.py)
            res = run_experts_sharded(eng, spec, table, table_cols, obs_col, coords_col, refs, ref_cols,
                                      self.local_select, pred_table=pred_tab, pred_cols=pred_cols,
                                      max_dist=self.pred_max_dist, optimise=optimise, predict=predict,
                                      min_obs=min_obs, theta_init=theta_init)
            dt = time.perf_counter() - t0
            self._shape_tables(pieces, res, sub, members, ok_load, dt, optimise, predict, model_name, dev_name,
                               config_id, D)
        tables = {}
        for name, lst in pieces.items():
            # rows back in the order the sequential loop would have appended them
            df = pd.concat([t[1] for t in lst], axis=0)
            if len(lst) > 1:
                df = df.iloc[np.argsort(df["_pos_"].values, kind="stable")]
            tables[f"{name}{table_suffix}"] = df.drop(columns="_pos_")
        # expert_locs + oi_config bookkeeping (local_experts.py:873-903; utils.py:1136-1273)
        tables[f"expert_locs{table_suffix}"] = self.expert_locs.set_index(coords_col)
        tables[f"oi_config{table_suffix}"] = pd.DataFrame(
            {"idx": [config_id], "datetime": [datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")],
             "config": [json.dumps(self.config, default=_json_default)]}).set_index("idx", drop=False)
        import torch.distributed as dist
        is_writer = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
        if store_path is not None and is_writer:
            self._write(store_path, tables)
        print(f"'run': {time.perf_counter() - t_run:.3f} seconds")
        return tables if return_tables else None

    @classmethod
    def run_from(cls, ref_oi, **run_kwargs):
        """Batched run of an already configured reference ``GPSat.local_experts.LocalExpertOI`` instance
        (the hook shown in INTEGRATION.md): its captured config dicts rebuild the driver."""
        cfg = ref_oi.config
        oi = cls(expert_loc_config=cfg.get("locations") or cfg.get("local_expert_locations"),
                 data_config=cfg.get("data"), model_config=cfg.get("model"), pred_loc_config=cfg.get("pred_loc"))
        if getattr(ref_oi, "expert_locs", None) is not None:
            oi.expert_locs = ref_oi.expert_locs
        return oi.run(**run_kwargs)

    def _device_name(self):
        import torch
        return torch.cuda.get_device_name(self.device)

    def _shape_tables(self, pieces, res, sub, members, ok_load, dt, optimise, predict, model_name, dev_name,
                      config_id, D):
        """Vectorised dict_of_array_to_table (local_experts.py:691-747) for a whole batch."""
        coords_col = self.coords_col
        E = len(sub)
        ref = sub[coords_col].values
        num_obs = res["num_obs"]
        too_few = res["too_few"]
        valid_idx = res.get("valid_idx", np.zeros(0, dtype=np.int64))
        keep = np.zeros(E, dtype=bool)
        keep[valid_idx] = True
        # experts whose parameters could not be loaded are skipped (local_experts.py:1099-1101)
        keep &= ok_load
        recorded = keep | too_few
        pos = np.asarray(members)
        Ev = len(valid_idx)
        vpos = np.full(E, -1)
        vpos[valid_idx] = np.arange(Ev)

        def midx(rows_ref):
            if len(coords_col) == 1:
                return pd.Index(rows_ref[:, 0], name=coords_col[0])
            return pd.MultiIndex.from_arrays([rows_ref[:, j] for j in range(len(coords_col))], names=coords_col)

        # run_details
        r = np.flatnonzero(recorded)
        fobj = np.full(E, np.nan)
        succ = np.zeros(E, dtype=bool)
        if Ev:
            fobj[valid_idx] = res["fobj"]
            if optimise:
                succ[valid_idx] = np.isin(res["status"], (1, 2))
        rt = np.where(keep, dt / max(int(keep.sum()), 1), np.nan)
        rd = pd.DataFrame({"_dim_0": 0, "num_obs": num_obs[r], "run_time": rt[r],
                           "objective_value": np.where(keep[r], fobj[r], np.nan),
                           "parameters_optimised": optimise, "optimise_success": succ[r] & keep[r],
                           "model": model_name, "device": np.where(keep[r], dev_name, ""),
                           "config_id": config_id, "_pos_": pos[r]}, index=midx(ref[r]))
        if len(r):
            pieces.setdefault("run_details", []).append((pos[r[0]], rd))
        k = np.flatnonzero(keep)
        if len(k) == 0:
            return
        first = pos[k[0]]
        same_table = False
        lp = self.load_params_config
        if lp is not None and not optimise:
            same_table = lp.get("file") is None
        # hyper-parameter tables (skipped when loading from and writing to the same table without optimising)
        if not same_table:
            th = res["theta"][vpos[k]]
            names = self.params_to_store or PARAM_NAMES
            if "lengthscales" in names:
                pieces.setdefault("lengthscales", []).append((first, pd.DataFrame(
                    {"_dim_0": np.tile(np.arange(D), len(k)), "lengthscales": th[:, :D].ravel(),
                     "_pos_": np.repeat(pos[k], D)},
                    index=midx(np.repeat(ref[k], D, axis=0)))))
            if "kernel_variance" in names:
                pieces.setdefault("kernel_variance", []).append((first, pd.DataFrame(
                    {"_dim_0": 0, "kernel_variance": th[:, D], "_pos_": pos[k]}, index=midx(ref[k]))))
            if "likelihood_variance" in names:
                pieces.setdefault("likelihood_variance", []).append((first, pd.DataFrame(
                    {"_dim_0": 0, "likelihood_variance": th[:, D + 1], "_pos_": pos[k]}, index=midx(ref[k]))))
            if "inducing_points" in res and (self.params_to_store is None or "inducing_points" in names):
                zo = res["z_offsets"]
                zc = res["inducing_points"]
                mk = np.diff(zo)[vpos[k]]
                rows = np.concatenate([np.arange(zo[v], zo[v + 1]) for v in vpos[k]])
                d0 = np.concatenate([np.repeat(np.arange(m), D) for m in mk])
                pieces.setdefault("inducing_points", []).append((first, pd.DataFrame(
                    {"_dim_0": d0, "_dim_1": np.tile(np.arange(D), len(rows)), "inducing_points": zc[rows].ravel(),
                     "_pos_": np.repeat(pos[k], mk * D)}, index=midx(np.repeat(ref[k], mk * D, axis=0)))))
        # preds
        if predict:
            poff = res["pred_offsets"]
            cnt = np.diff(poff)
            sel = np.concatenate([np.arange(poff[v], poff[v + 1]) for v in vpos[k]]) if len(k) else np.zeros(0, int)
            cntk = cnt[vpos[k]]
            dim0 = np.concatenate([np.arange(c) for c in cntk]) if len(k) else np.zeros(0, int)
            om = res["obs_mean"][vpos[k]]
            pr = {"_dim_0": dim0, "f*": res["fmean"][sel], "f*_var": res["fvar"][sel], "y_var": res["yvar"][sel],
                  "f_bar": np.repeat(om, cntk)}
            for ci, c in enumerate(coords_col):
                pr[f"pred_loc_{c}"] = res["pred_coords"][sel, ci]
            pr["_pos_"] = np.repeat(pos[k], cntk)
            pieces.setdefault("preds", []).append((first, pd.DataFrame(pr, index=midx(np.repeat(ref[k], cntk, axis=0)))))

    @staticmethod
    def _write(store_path, tables):
        os.makedirs(os.path.dirname(os.path.abspath(store_path)), exist_ok=True)
        with pd.HDFStore(store_path, mode="a") as store:     # raises ImportError without PyTables
            for k, v in tables.items():
                if k.startswith("expert_locs") and f"/{k}" in store.keys():
                    continue
                min_itemsize = {c: 64 for c in v.columns if c in ["model", "device"]}
                try:
                    if k.startswith("oi_config"):
                        store.append(key=k, value=v, min_itemsize={"config": 65536}, data_columns=["idx"])
                    elif k.startswith("expert_locs"):
                        store.append(key=k, value=v, data_columns=True)
                    else:
                        store.append(key=k, value=v, min_itemsize=min_itemsize)
                except ValueError as e:
                    print(e)


def get_results_from_tables(tables, table_suffix=""):
    """Convenience: strip the suffix so callers can index tables like the reference's
    get_results_from_h5file output (local_experts.py:1467-1620)."""
    return {re.sub(f"{re.escape(table_suffix)}$", "", k): v for k, v in tables.items()}
