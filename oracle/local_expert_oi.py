"""Sequential per-expert loop oracle.  TEST INFRASTRUCTURE ONLY.

Restates the body of LocalExpertOI.run (GPSat/local_experts.py:930-1260, rows O1/O2 of
SURVEY.md section 8a) on in-memory DataFrames, producing the tables the reference would
hand to pd.HDFStore.append (local_experts.py:538-543): ``run_details``, ``preds`` and one
table per hyper-parameter, each indexed by the expert's coords (local_experts.py:691-747).
The reference's LocalExpertOI itself cannot be imported here (PyTables/xarray/gpflow absent).
"""
from __future__ import annotations

import time

import numpy as np
import pandas as pd

from .gpr import OracleGPRModel
from . import selection as sel


def array_to_dataframe(x, name):
    """GPSat/utils.py:1437-1495 with reset_index=True."""
    if isinstance(x, (int, float, bool, str, np.floating, np.integer, np.bool_)):
        x = np.array([x])
    x = np.asarray(x)
    midx = pd.MultiIndex.from_product([np.arange(i) for i in x.shape],
                                      names=[f"_dim_{i}" for i in range(x.ndim)])
    return pd.DataFrame(x.flat, index=midx, columns=[name])


def dict_of_array_to_table(x, ref_loc: dict, concat=False, table=None):
    """local_experts.py:691-747 (+ utils.py:1619-1725)."""
    if len(x) == 0:
        return {}
    if concat:
        dfs = {}
        for k, v in x.items():
            nd = 1 if np.ndim(v) == 0 else np.ndim(v)
            dfs.setdefault(nd, []).append(array_to_dataframe(v, k))
        dfs = {k: pd.concat(v, join="outer", axis=1).reset_index() for k, v in dfs.items()}
    else:
        dfs = {k: array_to_dataframe(v, k).reset_index() for k, v in x.items()}
    names = list(ref_loc.keys())
    tup = tuple(ref_loc.values())
    for k, df in dfs.items():
        if len(names) == 1:
            df.index = pd.Index([tup[0]] * len(df), name=names[0])
        else:
            df.index = pd.MultiIndex.from_tuples([tup] * len(df), names=names)
    if not concat:
        return dfs
    return {(table if k == 1 else f"{table}_{k}"): v for k, v in dfs.items()}


def run_local_expert_oi(expert_locs: pd.DataFrame, data: dict, model: dict, pred_loc: dict,
                        optimise=True, predict=True, min_obs=3, model_cls=OracleGPRModel,
                        load_params=None, optimiser="scipy"):
    """Returns (tables: dict[str, DataFrame], per_expert: list[dict]) for the whole expert list.

    ``load_params``: optional callable(expert_row_dict) -> dict of parameter values (the
    predict-only path, local_experts.py:1075-1101, with the HDF5 lookup replaced by a callable).
    """
    coords_col = data["coords_col"]
    obs_col = data["obs_col"]
    src = data["data_source"]
    local_select = data["local_select"]
    global_select = data.get("global_select", [])
    init_params = dict(model.get("init_params", {}))
    constraints = model.get("constraints", None)
    optim_kwargs = dict(model.get("optim_kwargs", {}))
    pred_kwargs = dict(model.get("pred_kwargs", {}))

    store = {}
    per_expert = []
    src_cols = {c: src[c].values for c in src.columns}
    pred_cols = None
    if pred_loc.get("method", "expert_loc") == "from_dataframe":
        pdf = pred_loc["df"]
        pred_cols = {c: pdf[c].values for c in pdf.columns}

    def _append(save):
        for k, v in save.items():
            store.setdefault(k, []).append(v)

    for idx in range(len(expert_locs)):
        rl = expert_locs.iloc[[idx], :]
        rld = rl.iloc[0, :].to_dict()
        ref = {c: rld[c] for c in coords_col}
        t0 = time.time()
        # prediction locations (local_experts.py:958-965)
        if pred_cols is not None:
            pcoords, _ = sel.prediction_locations(pred_cols, coords_col, rld,
                                                  pred_loc.get("max_dist", None))
        else:
            pcoords = rl[coords_col].values.astype(np.float64)
        if len(pcoords) == 0:
            continue
        # global + local selection (local_experts.py:971-984)
        where = sel.expand_where_list(global_select, local_select, rld)
        gmask = sel.where_mask(src_cols, where)
        gcols = {c: v[gmask] for c, v in src_cols.items()}
        lmask = sel.local_select_mask(gcols, rld, local_select)
        df_local = pd.DataFrame({c: v[lmask] for c, v in gcols.items()})
        sel_idx = np.flatnonzero(gmask)[lmask]
        if len(df_local) < min_obs:
            rd = {"num_obs": len(df_local), "run_time": np.nan, "objective_value": np.nan,
                  "parameters_optimised": optimise, "optimise_success": False,
                  "model": model_cls.__name__[:64], "device": "", "config_id": 0}
            _append(dict_of_array_to_table(rd, ref, concat=True, table="run_details"))
            per_expert.append({"idx": idx, "sel_idx": sel_idx, "skipped": True})
            continue
        m = model_cls(data=df_local, obs_col=obs_col, coords_col=coords_col,
                      expert_loc=rl[coords_col].to_numpy().squeeze(), **init_params)
        if load_params is not None:
            m.set_parameters(**load_params(rld))
        if constraints is not None:
            cons = {k: dict(v) for k, v in constraints.items()}
            if init_params.get("coords_scale", None) is not None and "lengthscales" in cons:
                cons["lengthscales"]["scale"] = True
            m.set_parameter_constraints(cons, move_within_tol=True, tol=1e-2)
        if optimise:
            kw = dict(optim_kwargs)
            if model_cls is OracleGPRModel:
                kw["optimiser"] = optimiser
            ok = m.optimise_parameters(**kw)
        else:
            ok = False
        fobj = m.get_objective_function_value()
        hypes = m.get_parameters()
        if predict:
            pred = m.predict(coords=pcoords, **pred_kwargs)
            for ci, c in enumerate(coords_col):
                pred[f"pred_loc_{c}"] = pcoords[:, ci]
        else:
            pred = {}
        rd = {"num_obs": len(df_local), "run_time": time.time() - t0, "objective_value": fobj,
              "parameters_optimised": optimise, "optimise_success": ok,
              "model": model_cls.__name__[:64], "device": (m.cpu_name if m.gpu_name is None else m.gpu_name)[:64],
              "config_id": 0}
        save = {**dict_of_array_to_table(rd, ref, concat=True, table="run_details"),
                **dict_of_array_to_table(pred, ref, concat=True, table="preds"),
                **dict_of_array_to_table(hypes, ref, concat=False)}
        _append(save)
        per_expert.append({"idx": idx, "sel_idx": sel_idx, "skipped": False, "hypes": hypes,
                           "objective": fobj, "success": ok, "pred": pred,
                           "nfev": getattr(getattr(m, "opt_result", None), "nfev", None)
                           if not isinstance(getattr(m, "opt_result", None), dict)
                           else m.opt_result["nfev"]})
    tables = {k: pd.concat(v, axis=0) for k, v in store.items()}
    return tables, per_expert
