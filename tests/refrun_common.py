"""Shared pieces of the reference-run parity tests (CPU: tests/test_refrun_cpu.py, GPU: tests/test_gpu_refrun.py).

The fixture under tests/golden/refrun/ was produced by tests/golden/make_golden_refrun.py, which drives the UNMODIFIED
reference ``LocalExpertOI.run`` (GPSat/local_experts.py:761-1279) on configs/example_local_expert_oi.json with only
the file paths and the model name changed, and records every table and every ``HDFStore.append`` call.  The tests
re-run the same config through gpsat_b200's driver -- again changing nothing but paths and ``oi_model`` -- and compare
what lands in the (fake) store.
"""
import copy
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "refrun")
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)

import fake_hdfstore as fh  # noqa: E402
from make_golden_refrun import frame_from_json  # noqa: E402

VOLATILE = ("run_time", "model", "device", "datetime", "config")      # differ between any two runs / model classes
HYPERS = ("lengthscales", "kernel_variance", "likelihood_variance")


def golden(name):
    with open(os.path.join(GOLD, f"{name}.json")) as f:
        d = json.load(f)
    return {k: frame_from_json(v) for k, v in d["tables"].items()}, d.get("appends")


def setup_files(tmp, which="config.json", oi_model="B200GPRModel"):
    """Copy the fixture inputs to ``tmp``, put the observation table into a (fake) .h5 file, and return the example
    config with paths pointed there and ``oi_model`` replaced -- the one-line switch BASELINE.json asks for."""
    tmp = str(tmp)
    for nm in ("locations.csv", "2d_xy_grid.csv"):
        shutil.copy(os.path.join(GOLD, nm), os.path.join(tmp, nm))
    with open(os.path.join(GOLD, "data.json")) as f:
        data = frame_from_json(json.load(f))
    with pd.HDFStore(os.path.join(tmp, "ABC_binned.h5"), mode="a") as st:
        if "data" not in st:
            st.append("data", data, data_columns=True)
    with open(os.path.join(GOLD, which)) as f:
        cfg = json.load(f)
    cfg["results"]["dir"] = tmp
    for sec, key in (("locations", "source"), ("data", "data_source"), ("pred_loc", "df_file")):
        cfg[sec][key] = os.path.join(tmp, cfg[sec][key])
    if (cfg["model"].get("load_params") or {}).get("file"):
        cfg["model"]["load_params"]["file"] = os.path.join(tmp, cfg["model"]["load_params"]["file"])
    cfg["model"]["oi_model"] = oi_model
    return cfg, data, os.path.join(tmp, cfg["results"]["file"])


def make_oi(cls, cfg):
    return cls(expert_loc_config=copy.deepcopy(cfg["locations"]), data_config=copy.deepcopy(cfg["data"]),
               model_config=copy.deepcopy(cfg["model"]), pred_loc_config=copy.deepcopy(cfg["pred_loc"]))


def seed_store_with(store_path, tabs, names):
    with pd.HDFStore(store_path, mode="a") as st:
        for nm in names:
            st.append(nm, tabs[nm])
    fh.appends(store_path).clear()


def compare_structure(got: pd.DataFrame, ref: pd.DataFrame, name):
    assert list(got.index.names) == list(ref.index.names), (name, got.index.names, ref.index.names)
    assert list(got.columns) == list(ref.columns), (name, list(got.columns), list(ref.columns))
    assert len(got) == len(ref), (name, len(got), len(ref))
    unnamed = list(ref.index.names) == [None]        # an unnamed index is not part of the fixture
    g, r = got.reset_index(drop=unnamed), ref.reset_index(drop=unnamed)
    for c in r.columns:
        if c in VOLATILE:
            continue
        kg, kr = g[c].dtype.kind, r[c].dtype.kind
        assert kg == kr or {kg, kr} <= {"O", "U", "T"}, (name, c, g[c].dtype, r[c].dtype)
        if kr in "iub" or kr == "M" or c.startswith("pred_loc_") or c in ("f_bar", "x", "y", "t", "lon", "lat"):
            assert (g[c].values == r[c].values).all(), (name, c)


def compare_store(got, ref, optimised, rtol_pred, lml_slack=1e-6, suffix="", rtol_var=None):
    """got / ref: {table: DataFrame}.  Structure, order and integer / index content exactly; floats per BASELINE.json."""
    assert set(got) == set(ref), (sorted(got), sorted(ref))
    for name in ref:
        compare_structure(got[name], ref[name], name)
    rd, rrd = got[f"run_details{suffix}"], ref[f"run_details{suffix}"]
    f, fr = rd["objective_value"].values, rrd["objective_value"].values
    ran = ~np.isnan(fr)
    assert (np.isnan(f) == ~ran).all()
    if optimised:
        assert (f[ran] <= fr[ran] + lml_slack * np.abs(fr[ran])).all(), (f - fr) / np.abs(fr)
    else:
        np.testing.assert_allclose(f[ran], fr[ran], rtol=rtol_pred)
    assert (rd["optimise_success"].values == rrd["optimise_success"].values).all()
    assert (rd["device"].values[~ran] == "").all() and (rd["device"].values[ran] != "").all()
    p, pr = got[f"preds{suffix}"], ref[f"preds{suffix}"]
    starts = np.flatnonzero(np.r_[True, pr["_dim_0"].values[1:] == 0])
    ends = np.r_[starts[1:], len(pr)]
    for c in ("f*", "f*_var", "y_var"):
        a, b = p[c].values, pr[c].values
        tol = rtol_pred if (c == "f*" or rtol_var is None) else rtol_var
        for s, e in zip(starts, ends):
            np.testing.assert_allclose(a[s:e], b[s:e], rtol=tol, atol=tol * np.abs(b[s:e]).max(),
                                       err_msg=f"preds.{c}")
    for nm in HYPERS:
        if f"{nm}{suffix}" in ref and f"{nm}{suffix}" in got:
            a, b = got[f"{nm}{suffix}"][nm].values, ref[f"{nm}{suffix}"][nm].values
            # optimised: the LML is flat along some directions; the parameters are compared loosely, the fit tightly
            np.testing.assert_allclose(a, b, rtol=2e-2 if optimised else 1e-12, err_msg=nm)


def compare_appends(got, ref):
    """Same sequence of HDFStore.append calls: (table, keyword arguments, rows)."""
    norm = lambda apps: [(k, json.dumps(kw, sort_keys=True, default=str), int(n)) for k, kw, n in apps]
    assert norm(got) == norm(ref), "\n".join(f"{a}\n{b}" for a, b in zip(norm(got), norm(ref)) if a != b) + \
        f"\n(len {len(got)} vs {len(ref)})"


# ---------------------------------------------------------------------------------------------
# CPU stand-in for the engine call (tests only): the oracle's sequential loop behind run_experts_sharded's
# signature, so the driver's host logic (chunking, store traffic, resume, parameter loading) runs without a GPU
# ---------------------------------------------------------------------------------------------
def oracle_backed_sharded(eng, spec, table, table_cols, obs_col, coords_col, experts, ref_cols, local_select,
                          pred_table=None, pred_cols=None, max_dist=None, optimise=True, predict=True, min_obs=3,
                          theta_init=None, count_only=False, **kw):
    from oracle.local_expert_oi import run_local_expert_oi
    df = pd.DataFrame(np.asarray(table).T, columns=table_cols)
    eloc = pd.DataFrame(np.asarray(experts), columns=ref_cols)
    pred = {"method": "expert_loc"}
    if pred_table is not None:
        pred = {"method": "from_dataframe", "df": pd.DataFrame(np.asarray(pred_table).T, columns=pred_cols),
                "max_dist": max_dist}
    model = {"init_params": {"coords_scale": spec.coords_scale, "obs_mean": spec.obs_mean}, "constraints": spec.constraints}
    model["init_params"] = {k: v for k, v in model["init_params"].items() if v is not None}
    D = len(coords_col)
    lp = None
    if theta_init is not None:
        key = {tuple(r): k for k, r in enumerate(np.asarray(experts)[:, [ref_cols.index(c) for c in coords_col]])}
        lp = lambda row: (lambda k: {"lengthscales": theta_init[k, :D], "kernel_variance": theta_init[k, D],
                                     "likelihood_variance": theta_init[k, D + 1]})(key[tuple(row[c] for c in coords_col)])
    _, per = run_local_expert_oi(eloc, {"data_source": df, "obs_col": obs_col, "coords_col": coords_col,
                                        "local_select": local_select}, model, pred, optimise=optimise,
                                 predict=predict, min_obs=min_obs, load_params=lp)
    E = len(eloc)
    num_obs = np.zeros(E, dtype=np.int64)
    has_pred, too_few = np.zeros(E, dtype=bool), np.zeros(E, dtype=bool)
    live = []
    for pe in per:
        has_pred[pe["idx"]] = True
        num_obs[pe["idx"]] = len(pe["sel_idx"])
        too_few[pe["idx"]] = pe["skipped"]
        if not pe["skipped"]:
            live.append(pe)
    # experts without prediction locations never reach the selection in the sequential loop: count them here
    from oracle import selection as sel
    cols = {c: df[c].values for c in df.columns}
    for i in np.flatnonzero(~has_pred):
        num_obs[i] = int(sel.local_select_mask(cols, eloc.iloc[i].to_dict(), local_select).sum())
    valid = has_pred & ~too_few
    out = dict(num_obs=num_obs, has_pred=has_pred, too_few=too_few, valid=valid, valid_idx=np.flatnonzero(valid),
               n_valid=int(valid.sum()))
    if count_only or not live:
        return out
    out["theta"] = np.array([np.r_[pe["hypes"]["lengthscales"], pe["hypes"]["kernel_variance"],
                                   pe["hypes"]["likelihood_variance"]] for pe in live])
    out["fobj"] = np.array([pe["objective"] for pe in live])
    out["status"] = np.array([1 if pe["success"] else 5 for pe in live], dtype=np.int32)
    if predict:
        cnt = [len(pe["pred"]["f*"]) for pe in live]
        out["pred_offsets"] = np.r_[0, np.cumsum(cnt)].astype(np.int64)
        out["obs_mean"] = np.array([pe["pred"]["f_bar"][0] for pe in live])
        out["pred_coords"] = np.concatenate([np.column_stack([pe["pred"][f"pred_loc_{c}"] for c in coords_col])
                                             for pe in live])
        for k, src in (("fmean", "f*"), ("fvar", "f*_var"), ("yvar", "y_var")):
            out[k] = np.concatenate([pe["pred"][src] for pe in live])
    return out
