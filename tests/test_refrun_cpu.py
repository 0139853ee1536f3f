"""CPU: the drop-in boundary against the UNMODIFIED reference orchestrator's recorded behaviour.

tests/golden/refrun/ holds what GPSat's own ``LocalExpertOI.run`` (local_experts.py:761-1279) wrote to its HDF5 store
when driven on configs/example_local_expert_oi.json (generator: tests/golden/make_golden_refrun.py).  Here, without a
GPU:

  * the oracle's sequential loop and table shapers reproduce those tables (pins oracle/local_expert_oi.py -- rows O1 /
    O2 and GPSat/utils.py:1437-1495,1619-1725 -- to the reference's own output);
  * gpsat_b200's config ingestion (set_expert_locations -> dataloader.load with add_data_to_col / col_funcs /
    row_select / col_select / sort_by) yields the reference's expert-location table from the unchanged JSON;
  * gpsat_b200.local_experts.LocalExpertOI.run -- with the engine call replaced by an oracle-backed stand-in, everything
    else as shipped -- produces the same tables through the same sequence of HDFStore.append calls (same keyword
    arguments, same row counts), resumes an interrupted run from ``run_details``, refuses a changed config, and runs
    the predict-only second pass from the ``_SMOOTHED`` parameter tables of the same file;
  * get_results_from_h5file returns the reference reader's frames.

The GPU counterpart (tests/test_gpu_refrun.py) repeats the store comparison with the real CUDA engine.
"""
import numpy as np
import pandas as pd
import pytest

import refrun_common as rc
from refrun_common import fh


@pytest.fixture()
def store(tmp_path):
    fh.install()
    yield tmp_path
    fh.uninstall()


@pytest.fixture()
def cpu_driver(monkeypatch):
    """the shipped driver with ONLY the engine call swapped for the oracle (no GPU in this container)"""
    from gpsat_b200 import local_experts as le
    import gpsat_b200
    monkeypatch.setattr(le, "run_experts_sharded", rc.oracle_backed_sharded)
    monkeypatch.setattr(gpsat_b200, "get_engine", lambda device=0: None)
    monkeypatch.setattr(le.LocalExpertOI, "_device_name", lambda self: "oracle-cpu")
    return le


def test_oracle_loop_reproduces_the_reference_run(store):
    """oracle/local_expert_oi.py (the checker of every GPU parity test) == GPSat's LocalExpertOI.run output"""
    from oracle.local_expert_oi import run_local_expert_oi
    cfg, data, _ = rc.setup_files(store)
    ref, _ = rc.golden("scenario_a")
    eloc = ref["expert_locs"].reset_index()
    pred = {"method": "from_dataframe", "df": pd.read_csv(cfg["pred_loc"]["df_file"]),
            "max_dist": cfg["pred_loc"]["max_dist"]}
    dcfg = {"data_source": data, "obs_col": cfg["data"]["obs_col"], "coords_col": cfg["data"]["coords_col"],
            "local_select": cfg["data"]["local_select"], "global_select": cfg["data"]["global_select"]}
    tabs, per = run_local_expert_oi(eloc, dcfg, {k: v for k, v in cfg["model"].items() if k != "oi_model"}, pred)
    for nm in ("run_details", "preds") + rc.HYPERS:
        rc.compare_structure(tabs[nm].drop(columns=["config_id"], errors="ignore"),
                             ref[nm].drop(columns=["config_id"], errors="ignore"), nm)
    ran = ~np.isnan(ref["run_details"]["objective_value"].values)
    np.testing.assert_allclose(tabs["run_details"]["objective_value"].values[ran],
                               ref["run_details"]["objective_value"].values[ran], rtol=1e-9)
    for c in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(tabs["preds"][c].values, ref["preds"][c].values, rtol=1e-6)
    for nm in rc.HYPERS:
        np.testing.assert_allclose(tabs[nm][nm].values, ref[nm][nm].values, rtol=1e-5)


def test_example_config_locations_are_ingested_like_the_reference(store):
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, _ = rc.setup_files(store)
    oi = LocalExpertOI(expert_loc_config=cfg["locations"])
    ref, _ = rc.golden("scenario_a")
    want = ref["expert_locs"].reset_index()
    got = oi.expert_locs
    # 8 csv rows x 3 dates (add_data_to_col) -> date / t from col_funcs -> date == 2020-03-05 & lat >= 60 -> 7 rows
    assert list(got.columns) == ["x", "y", "t", "date", "lon", "lat"]
    assert len(got) == len(want) == 7
    for c in ("x", "y", "t", "lon", "lat"):
        np.testing.assert_array_equal(got[c].values, want[c].values)
    assert (got["date"].values.astype("datetime64[D]") == np.datetime64("2020-03-05")).all()
    assert got["t"].dtype == np.float64 and (got["t"] == 18326.0).all()
    # the config as captured is what check_prev_oi_config compares between runs
    assert oi.config["locations"]["add_data_to_col"] == {"date": ["2020-03-04", "2020-03-05", "2020-03-06"]}
    assert oi.config["locations"]["sort_by"] == "date"


def test_dataloader_pieces():
    from gpsat_b200 import dataloader as dl
    df = pd.DataFrame({"A": [1, 2, 3], "B": [4, 5, 6]})
    # the reference's docstring examples (dataloader.py:1458-1477, 1563-1569; utils.py:382-402)
    out = dl.add_data_to_col(df, {"C": [7, 8]})
    assert out["C"].tolist() == [7, 7, 7, 8, 8, 8] and out.index.tolist() == [0, 1, 2, 0, 1, 2]
    assert len(dl.add_data_to_col(df, {"a": [1, 2, 3, 4], "b": [5, 6, 7, 8]})) == 48
    assert dl.load(df, where={"col": "A", "comp": ">=", "val": 2})["B"].tolist() == [5, 6]
    assert dl.config_func(func="lambda x, y: x + y", args=[1, 1]) == 2
    assert dl.config_func(func="==", args=[1, 1]) is True
    assert dl.config_func(func="<=", col_args=["A", "B"], df=df).tolist() == [True, True, True]
    assert dl.config_func(func="cumprod", source="numpy", df=df, kwargs={"axis": 0},
                          col_args=[["A", "B"]]).tolist() == [[1, 4], [2, 20], [6, 120]]
    d2 = df.copy()
    dl.add_cols(d2, {"C": {"func": "lambda x: x + 1", "col_args": "A"}, ("D", "E"): {"func": "lambda x: (x, -x)", "col_args": "B"}})
    assert d2["C"].tolist() == [2, 3, 4] and d2["E"].tolist() == [-4, -5, -6]
    sel = dl.load(df, row_select=[{"col": "A", "comp": ">", "val": 1}, {"col": "B", "comp": "<", "val": 6, "negate": True}],
                  col_select=["B"])
    assert sel["B"].tolist() == [6] and list(sel.columns) == ["B"]
    with pytest.raises(AssertionError):
        dl.load(df, col_select=["nope"])
    with pytest.raises(AssertionError):
        dl.where_mask(df, {"col": "A", "comp": "~", "val": 1})
    w = dl.get_where_list([{"col": "lat", "comp": ">=", "val": 60},
                           {"loc_col": "t", "src_col": "date",
                            "func": "lambda x,y: np.datetime64(pd.to_datetime(x+y, unit='D'))"}],
                          local_select=[{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
                                        {"col": ["x", "y"], "comp": "<", "val": 3e5}], ref_loc={"t": 18326.0})
    assert [x["comp"] for x in w] == [">=", "<=", ">="] and w[1]["val"] == np.datetime64("2020-03-09")
    assert dl.store_where(w[0]) == "lat>=60" and dl.store_where(w[2]).startswith('date>="2020-03-01')


def test_driver_store_traffic_matches_the_reference_run(store, cpu_driver):
    """scenario A: optimise + predict, interrupted after 3 locations and resumed -- same tables, same appends"""
    cfg, _, store_path = rc.setup_files(store)
    oi = rc.make_oi(cpu_driver.LocalExpertOI, cfg)
    full = oi.expert_locs.copy(True)
    oi.expert_locs = full.iloc[:3].copy(True)
    rk = dict(cfg["run_kwargs"], store_every=2, max_batch=2)      # max_batch is the flush unit (= store_every here)
    assert oi.run(store_path=store_path, **rk) is None
    ref1, app1 = rc.golden("scenario_a_part1")
    got1 = {k: v for k, v in fh.tables(store_path).items()}
    rc.compare_store(got1, ref1, optimised=True, rtol_pred=1e-4)
    rc.compare_appends(fh.appends(store_path), app1)
    # second half: a fresh driver on the same config finds the first three in run_details and skips them
    oi = rc.make_oi(cpu_driver.LocalExpertOI, cfg)
    oi.run(store_path=store_path, **rk)
    ref, app = rc.golden("scenario_a")
    rc.compare_store(dict(fh.tables(store_path)), ref, optimised=True, rtol_pred=1e-4)
    rc.compare_appends(fh.appends(store_path), app)
    assert (fh.tables(store_path)["run_details"]["config_id"] == 1).all()
    assert len(fh.tables(store_path)["oi_config"]) == 1            # same config -> no new row
    # a third run has nothing left to do and writes nothing
    n_app = len(fh.appends(store_path))
    rc.make_oi(cpu_driver.LocalExpertOI, cfg).run(store_path=store_path, **rk)
    assert [a[0] for a in fh.appends(store_path)[n_app:]] == ["expert_locs"] and fh.appends(store_path)[-1][2] == 0
    # a different config on the same store is refused when check_config_compatible is on ...
    cfg2, _, _ = rc.setup_files(store)
    cfg2["data"]["local_select"][2]["val"] = 250000
    with pytest.raises(AssertionError, match="did not have values that matched"):
        rc.make_oi(cpu_driver.LocalExpertOI, cfg2).run(store_path=store_path, **dict(rk, check_config_compatible=True))
    # ... and gets the next config id (a new oi_config row) when it is off, as in the example's run_kwargs
    assert len(fh.tables(store_path)["oi_config"]) == 2
    assert fh.tables(store_path)["oi_config"]["idx"].tolist() == [1, 2]


def test_driver_predict_only_from_smoothed_tables(store, cpu_driver):
    """scenario B (configs[1] of BASELINE.json): load_params from the _SMOOTHED tables of the same file, optimise=False"""
    cfg, _, store_path = rc.setup_files(store, "config_b.json")
    ref, app = rc.golden("scenario_b")
    before = [k for k in ref if not k.endswith("_SMOOTHED") or k in [f"{h}_SMOOTHED" for h in rc.HYPERS]]
    rc.seed_store_with(store_path, ref, before)
    oi = rc.make_oi(cpu_driver.LocalExpertOI, cfg)
    oi.run(store_path=store_path, **cfg["run_kwargs"])
    got = dict(fh.tables(store_path))
    rc.compare_store(got, ref, optimised=False, rtol_pred=1e-8, suffix="_SMOOTHED")
    # loading from and writing to the same suffix without optimising: the parameter tables are not appended again
    new = [a for a in app if a[0].endswith("_SMOOTHED") and a[0] not in [f"{h}_SMOOTHED" for h in rc.HYPERS]]
    rc.compare_appends(fh.appends(store_path), new)


def test_driver_previous_parameters_sequential_fallback(store, cpu_driver):
    """scenario C: load_params={"previous": True} makes experts order-dependent (EMA warm start,
    local_experts.py:1079-1083,1200-1217): the driver falls back to one expert per engine call in list order and
    flushes every store_every experts like the reference -- same tables, same appends."""
    cfg, _, store_path = rc.setup_files(store, "config_c.json")
    oi = rc.make_oi(cpu_driver.LocalExpertOI, cfg)
    oi.run(store_path=store_path, **dict(cfg["run_kwargs"], store_every=2))
    ref, app = rc.golden("scenario_c")
    rc.compare_store(dict(fh.tables(store_path)), ref, optimised=True, rtol_pred=1e-4)
    rc.compare_appends(fh.appends(store_path), app)


def test_load_params_forms_and_missing_experts(store, cpu_driver):
    """an expert missing from the parameter tables is skipped (not a crash), NaN parameters fall back to the
    model default, fixed values can be given inline, and index_adjust shifts the lookup key."""
    cfg, _, store_path = rc.setup_files(store, "config_b.json")
    ref, _ = rc.golden("scenario_b")
    ls = ref["lengthscales_SMOOTHED"].copy()
    kv = ref["kernel_variance_SMOOTHED"].copy()
    nv = ref["likelihood_variance_SMOOTHED"].copy()
    first = ls.index[0]
    ls, kv, nv = ls[ls.index != first], kv[kv.index != first], nv[nv.index != first]      # expert 0 is missing
    kv.iloc[1, kv.columns.get_loc("kernel_variance")] = np.nan                               # NaN -> default (1.0)
    oi = rc.make_oi(cpu_driver.LocalExpertOI, cfg)
    oi.load_params_config = {"file": {"lengthscales_S": ls, "kernel_variance_S": kv, "likelihood_variance_S": nv},
                             "table_suffix": "_S"}
    theta, ok = oi._load_theta(oi.expert_locs)
    assert ok.tolist() == [False, True, True, True, True, False, False]
    assert theta[2, 3] == 1.0 and theta[1, 3] == kv["kernel_variance"].iloc[0]
    np.testing.assert_array_equal(theta[1, :3], ls["lengthscales"].values[:3])
    tabs = oi.run(store_path=None, optimise=False, table_suffix="_X")
    rd = tabs["run_details_X"]
    assert len(rd) == 5 and first not in rd.index       # 4 ran + the far one with too few observations; expert 0 skipped
    assert rd["num_obs"].tolist()[-1] == 0
    # inline fixed values (load_params without a file)
    oi.load_params_config = {"lengthscales": [2.0, 3.0, 4.0], "likelihood_variance": 0.005}
    theta, ok = oi._load_theta(oi.expert_locs)
    assert ok.all() and theta[0].tolist() == [2.0, 3.0, 4.0, 1.0, 0.005]
    # index_adjust: look the parameters up one day earlier
    ls2 = ref["lengthscales_SMOOTHED"].reset_index()
    ls2["t"] -= 1.0
    oi.load_params_config = {"file": {"lengthscales": ls2.set_index(["x", "y", "t"])}, "param_names": ["lengthscales"],
                             "index_adjust": {"t": {"func": "lambda x: x - 1"}}}
    theta, ok = oi._load_theta(oi.expert_locs)
    assert ok.sum() == 5 and np.array_equal(theta[0, :3], ref["lengthscales_SMOOTHED"]["lengthscales"].values[:3])


def test_get_results_from_h5file_matches_the_reference_reader(store):
    import json
    import os
    from gpsat_b200.local_experts import get_results_from_h5file
    ref, _ = rc.golden("scenario_a")
    path = os.path.join(str(store), "results.h5")
    rc.seed_store_with(path, ref, list(ref))
    dfs, oi_config = get_results_from_h5file(path)
    with open(os.path.join(rc.GOLD, "results_a.json")) as f:
        want = json.load(f)
    assert len(oi_config) == want["n_config"] and sorted(oi_config[0]) == want["config_keys"]
    assert set(dfs) == set(want["tables"])
    for k, v in want["tables"].items():
        w = rc.frame_from_json(v)
        assert list(dfs[k].columns) == list(w.columns), k
        assert len(dfs[k]) == len(w)
        for c in w.columns:
            if w[c].dtype.kind == "f":
                np.testing.assert_array_equal(dfs[k][c].values, w[c].values, err_msg=f"{k}.{c}")
    sel, _ = get_results_from_h5file(path, select_tables=["preds", "run_details"], merge_on_expert_locations=False)
    assert set(sel) == {"preds", "run_details"} and "lon" not in sel["preds"].columns
