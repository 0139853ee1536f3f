"""Drive the UNMODIFIED reference orchestrator -- GPSat.local_experts.LocalExpertOI.run (local_experts.py:761-1279) --
end to end in the authoring container and freeze what it hands to the HDF5 store.

    python tests/golden/make_golden_refrun.py        -> tests/golden/refrun/  (inputs + tables as JSON)

What runs unmodified: LocalExpertOI.__init__/set_* (config capture, DataLoader.load with add_data_to_col /
col_funcs / row_select / col_select / sort_by), get_previous_oi_config / check_prev_oi_config, the expert_locs
bookkeeping, the resume anti-join, the per-expert loop (PredictionLocations, get_where_list + HDFStore.select
push-down, DataLoader.local_data_select), load_params/_read_params_from_file, dict_of_array_to_table and the
store_every flushes.  What is substituted, and how:

  * the model class: ``oi_model = {"path_to_model": "oracle.gpr", "model_name": "OracleGPRModel"}`` -- the reference's
    own custom-model hook (local_experts.py:319-325).  GPflowGPRModel needs gpflow/tensorflow (absent);
  * pandas.HDFStore: tests/fake_hdfstore.py (PyTables absent) -- records every append with its kwargs;
  * module-level imports that are absent (tests/golden/ref_stubs.py); the xarray stub carries a 12-line
    ``DataArray.from_series`` (an unstack) because _read_params_from_file routes parameter tables through it
    (dataloader.py:2585-2598).

The config is configs/example_local_expert_oi.json with the file paths pointed at synthetic files and oi_model
changed -- nothing else.  Scenario A: optimise + predict, interrupted and resumed.  Scenario B: predict-only from the
"_SMOOTHED" parameter tables of the same file (configs[1] of BASELINE.json).
"""
import copy
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "refrun")
WORK = "/tmp/gpsat_refrun"


# ---------------------------------------------------------------------------------------------
# synthetic inputs on disk (the reference's example blobs are not distributed: .MISSING_LARGE_BLOBS)
# ---------------------------------------------------------------------------------------------
def make_inputs(work):
    rng = np.random.default_rng(20200305)
    n = 5200
    x = rng.uniform(-6e5, 6e5, n)
    y = rng.uniform(-6e5, 6e5, n)
    t = rng.integers(18320, 18333, n).astype(np.float64)
    obs = 0.1 * np.sin(x / 2e5) + 0.05 * np.cos(y / 1.5e5) + 0.02 * np.sin((t - 18326) / 3) + rng.normal(0, 0.05, n)
    lat = 90.0 - np.hypot(x, y) / 111_000.0
    lon = np.degrees(np.arctan2(x, -y))
    data = pd.DataFrame({"x": x, "y": y, "t": t, "obs": obs, "lat": lat, "lon": lon,
                         "date": pd.to_datetime(t, unit="D")})
    # expert locations (x, y, lon, lat like data/locations/example_expert_locations_arctic_no_date.csv):
    # 5 ordinary, one with lat < 60 (dropped by row_select), one with no data nearby (recorded, not run),
    # one far from every prediction location (skipped silently)
    ex = np.array([0.0, 2e5, -2e5, 1e5, -3e5, 3.5e6, 5e6, 9e6])
    ey = np.array([0.0, -1e5, 2e5, 3e5, -2e5, 0.0, 5e6 * 0 + 4e5, 9e6])
    ex[6], ey[6] = 1.6e6, 1.6e6
    locs = pd.DataFrame({"x": ex, "y": ey, "lon": np.degrees(np.arctan2(ex, -ey)),
                         "lat": 90.0 - np.hypot(ex, ey) / 111_000.0})
    locs.loc[6, "lat"] = 61.0        # far from the data, kept by row_select
    locs.loc[7, "lat"] = 61.0        # far from the data AND from the prediction grid
    g = np.arange(-4.5e5, 4.5e5 + 1, 5e4)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    pred = pd.DataFrame({"x": np.r_[gx.ravel(), 1.6e6], "y": np.r_[gy.ravel(), 1.6e6]})
    os.makedirs(work, exist_ok=True)
    locs.to_csv(os.path.join(work, "locations.csv"), index=False)
    pred.to_csv(os.path.join(work, "2d_xy_grid.csv"), index=False)
    return data, locs, pred


def example_config(work, oi_model):
    """configs/example_local_expert_oi.json with only paths and the model name changed"""
    with open("/root/reference/configs/example_local_expert_oi.json") as f:
        cfg = json.load(f)
    cfg.pop("comment")
    cfg["results"] = {"dir": work, "file": "ABC_binned_oi.h5"}
    cfg["locations"]["source"] = os.path.join(work, "locations.csv")
    cfg["data"]["data_source"] = os.path.join(work, "ABC_binned.h5")
    cfg["pred_loc"]["df_file"] = os.path.join(work, "2d_xy_grid.csv")
    cfg["model"]["oi_model"] = oi_model
    return cfg


# ---------------------------------------------------------------------------------------------
# JSON (de)serialisation of DataFrames: exact float round trip, dtypes and index names kept
# ---------------------------------------------------------------------------------------------
def frame_to_json(df):
    flat = df.reset_index()
    cols = {}
    for c in flat.columns:
        v = flat[c]
        if np.issubdtype(v.dtype, np.datetime64):
            cols[c] = {"dtype": "datetime64[ns]", "values": [str(x) for x in v.values.astype("datetime64[ns]")]}
        elif v.dtype == bool:
            cols[c] = {"dtype": "bool", "values": [bool(x) for x in v.values]}
        elif np.issubdtype(v.dtype, np.integer):
            cols[c] = {"dtype": "int64", "values": [int(x) for x in v.values]}
        elif np.issubdtype(v.dtype, np.floating):
            cols[c] = {"dtype": "float64", "values": [None if np.isnan(x) else float(x) for x in v.values]}
        else:
            cols[c] = {"dtype": "str", "values": [str(x) for x in v.values]}
    return {"index_names": list(df.index.names), "columns": list(df.columns), "data": cols}


def frame_from_json(d):
    data = {}
    for c, v in d["data"].items():
        if v["dtype"] == "float64":
            data[c] = np.array([np.nan if x is None else x for x in v["values"]], dtype=np.float64)
        elif v["dtype"] == "datetime64[ns]":
            data[c] = np.array(v["values"], dtype="datetime64[ns]")
        elif v["dtype"] == "str":
            data[c] = np.array(v["values"], dtype=object)
        else:
            data[c] = np.array(v["values"], dtype=v["dtype"])
    flat = pd.DataFrame(data)
    if d["index_names"] != [None]:
        flat = flat.set_index(d["index_names"])
    return flat[d["columns"]]


def dump_store(path, name):
    import fake_hdfstore as fh
    tabs = {k: frame_to_json(v) for k, v in fh.tables(path).items() if k != "data"}
    apps = [[k, {a: (b if not isinstance(b, dict) else dict(b)) for a, b in kw.items()}, n]
            for k, kw, n in fh.appends(path) if k != "data"]
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump({"tables": tabs, "appends": apps}, f)
    return tabs


def main():
    from ref_stubs import install_stubs, _mod
    install_stubs(orchestrator=True)

    # the one functional stub: xarray.DataArray.from_series == unstack of a MultiIndex series into a dense array
    class DataArray:
        def __init__(self, data, coords=None):
            self.data = self.values = np.asarray(data)
            self.coords = coords or {}

        @classmethod
        def from_series(cls, s):
            levels = [np.unique(s.index.get_level_values(i)) for i in range(s.index.nlevels)]
            out = np.full([len(lv) for lv in levels], np.nan)
            pos = tuple(np.searchsorted(levels[i], s.index.get_level_values(i)) for i in range(s.index.nlevels))
            out[pos] = s.values
            return cls(out, coords={nm: lv for nm, lv in zip(s.index.names, levels)})

    import xarray
    xarray.DataArray = DataArray
    sys.modules["xarray.core.dataarray"].DataArray = DataArray

    # the reference predates pandas 3: its example col_funcs call ndarray.astype('datetime64[D]') on the .values of a
    # string column, which needs object-dtype strings (pandas < 3 behaviour)
    pd.set_option("future.infer_string", False)
    import fake_hdfstore as fh
    fh.install()
    shutil.rmtree(WORK, ignore_errors=True)
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    data, locs, pred = make_inputs(WORK)
    with pd.HDFStore(os.path.join(WORK, "ABC_binned.h5"), mode="a") as st:
        st.append("data", data, data_columns=True)
    for nm in ("locations.csv", "2d_xy_grid.csv"):
        shutil.copy(os.path.join(WORK, nm), os.path.join(OUT, nm))
    with open(os.path.join(OUT, "data.json"), "w") as f:
        json.dump(frame_to_json(data), f)

    from GPSat.local_experts import LocalExpertOI
    oi_model = {"path_to_model": "oracle.gpr", "model_name": "OracleGPRModel"}
    cfg = example_config(WORK, oi_model)
    with open(os.path.join(OUT, "config.json"), "w") as f:      # paths relative to the fixture directory
        c2 = copy.deepcopy(cfg)
        c2["results"]["dir"] = "."
        for sec, key in (("locations", "source"), ("data", "data_source"), ("pred_loc", "df_file")):
            c2[sec][key] = os.path.basename(c2[sec][key])
        json.dump(c2, f, indent=1)
    store_path = os.path.join(cfg["results"]["dir"], cfg["results"]["file"])

    # ---- scenario A, first half: the run is "interrupted" after the first 3 expert locations ----
    def make():
        return LocalExpertOI(expert_loc_config=copy.deepcopy(cfg["locations"]), data_config=copy.deepcopy(cfg["data"]),
                             model_config=copy.deepcopy(cfg["model"]), pred_loc_config=copy.deepcopy(cfg["pred_loc"]))

    oi = make()
    print(oi.expert_locs)
    full_locs = oi.expert_locs.copy(True)
    oi.expert_locs = full_locs.iloc[:3].copy(True)
    rk = dict(cfg["run_kwargs"], store_every=2)
    oi.run(store_path=store_path, **rk)
    dump_store(store_path, "scenario_a_part1")
    # ---- second half: same config, all locations -> the first three are found in run_details and skipped ----
    oi = make()
    oi.run(store_path=store_path, **rk)
    tabs = dump_store(store_path, "scenario_a")
    print({k: len(v["data"][v["columns"][0]]["values"]) for k, v in tabs.items()})
    # the reference's reader of that file (local_experts.py:1467-1620)
    from GPSat.local_experts import get_results_from_h5file
    dfs, oi_cfg = get_results_from_h5file(store_path)
    with open(os.path.join(OUT, "results_a.json"), "w") as f:
        json.dump({"tables": {k: frame_to_json(v) for k, v in dfs.items()}, "n_config": len(oi_cfg),
                   "config_keys": sorted(oi_cfg[0].keys())}, f)

    # ---- scenario B: "smoothed" parameter tables in the same file, predict-only (the example's second config) ----
    with pd.HDFStore(store_path, mode="a") as st:
        for nm, fac in (("lengthscales", 1.25), ("kernel_variance", 0.8), ("likelihood_variance", 1.1)):
            df = st.get(nm).copy()
            df[nm] = df[nm] * fac
            st.append(f"{nm}_SMOOTHED", df)
    cfg_b = copy.deepcopy(cfg)
    cfg_b["model"]["load_params"] = {"file": store_path, "table_suffix": "_SMOOTHED"}
    cfg_b["run_kwargs"].update(optimise=False, table_suffix="_SMOOTHED")
    oi = LocalExpertOI(expert_loc_config=cfg_b["locations"], data_config=cfg_b["data"], model_config=cfg_b["model"],
                       pred_loc_config=cfg_b["pred_loc"])
    oi.run(store_path=store_path, **cfg_b["run_kwargs"])
    dump_store(store_path, "scenario_b")
    with open(os.path.join(OUT, "config_b.json"), "w") as f:
        c2 = copy.deepcopy(cfg_b)
        c2["results"]["dir"] = "."
        c2["model"]["load_params"]["file"] = cfg["results"]["file"]
        for sec, key in (("locations", "source"), ("data", "data_source"), ("pred_loc", "df_file")):
            c2[sec][key] = os.path.basename(c2[sec][key])
        json.dump(c2, f, indent=1)
    # ---- scenario C: load_params={"previous": True} -- every expert starts from the EMA of its predecessors ----
    cfg_c = copy.deepcopy(cfg)
    cfg_c["results"]["file"] = "ABC_binned_oi_previous.h5"
    cfg_c["model"]["load_params"] = {"previous": True}
    store_c = os.path.join(cfg_c["results"]["dir"], cfg_c["results"]["file"])
    oi = LocalExpertOI(expert_loc_config=copy.deepcopy(cfg_c["locations"]), data_config=copy.deepcopy(cfg_c["data"]),
                       model_config=copy.deepcopy(cfg_c["model"]), pred_loc_config=copy.deepcopy(cfg_c["pred_loc"]))
    oi.run(store_path=store_c, **dict(cfg_c["run_kwargs"], store_every=2))
    dump_store(store_c, "scenario_c")
    with open(os.path.join(OUT, "config_c.json"), "w") as f:
        c2 = copy.deepcopy(cfg_c)
        c2["results"]["dir"] = "."
        for sec, key in (("locations", "source"), ("data", "data_source"), ("pred_loc", "df_file")):
            c2[sec][key] = os.path.basename(c2[sec][key])
        json.dump(c2, f, indent=1)
    fh.uninstall()


if __name__ == "__main__":
    main()
