"""GPU: the reference's whole local-expert workflow through this package, checked stage by stage against the oracle:
raw along-track points --bin_data_by--> observation table --LocalExpertOI.run(optimise)--> hyper-parameter tables
--smooth_hyperparameter_table--> *_SMOOTHED tables --LocalExpertOI.run(optimise=False, load_params)--> predictions
--glue_local_predictions_2d--> gridded field.   (examples/inline_example.py:150-534 is this chain.)"""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import binning as ob  # noqa: E402
from oracle import postproc as opp  # noqa: E402
from oracle.local_expert_oi import run_local_expert_oi  # noqa: E402


def test_bin_optimise_smooth_predict_glue():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpsat_b200 import build
    build.build()
    from gpsat_b200.dataprepper import DataPrep
    from gpsat_b200.local_experts import LocalExpertOI
    from gpsat_b200 import postprocessing as pp
    rng = np.random.default_rng(77)
    # --- raw points -> binned observation table (25 km bins, one table per day)
    n = 120_000
    raw = pd.DataFrame({"x": rng.uniform(-6e5, 6e5, n), "y": rng.uniform(-6e5, 6e5, n),
                        "t": rng.integers(18323, 18330, n).astype(float)})
    raw["z"] = 0.1 * np.sin(raw["x"] / 2e5) + 0.05 * np.cos(raw["y"] / 1.5e5) + rng.normal(0, 0.05, n)
    kw = dict(x_range=[-6e5, 6e5], y_range=[-6e5, 6e5], grid_res=25_000.0)
    binned = DataPrep.bin_data_by(raw, by_cols=["t"], val_col="z", **kw)
    ref_binned = ob.bin_data_by(raw, ["t"], "z", "x", "y", kw["x_range"], kw["y_range"], kw["grid_res"])
    assert binned.index.equals(ref_binned.index)
    np.testing.assert_allclose(binned["z"].values, ref_binned["z"].values, rtol=1e-12)
    obs = ref_binned.dropna().reset_index()           # both arms continue from the same table
    # --- optimise on a 3 x 3 expert lattice
    ex, ey = np.meshgrid([-2e5, 0.0, 2e5], [-2e5, 0.0, 2e5])
    eloc = pd.DataFrame({"x": ex.ravel(), "y": ey.ravel(), "t": 18326.0})
    gx, gy = np.meshgrid(np.arange(-3e5, 3e5 + 1, 5e4), np.arange(-3e5, 3e5 + 1, 5e4))
    ploc = pd.DataFrame({"x": gx.ravel(), "y": gy.ravel()})
    data = {"data_source": obs, "obs_col": "z", "coords_col": ["x", "y", "t"],
            "local_select": [{"col": "t", "comp": "<=", "val": 3}, {"col": "t", "comp": ">=", "val": -3},
                             {"col": ["x", "y"], "comp": "<", "val": 200_000}]}
    model = {"oi_model": "B200GPRModel", "init_params": {"coords_scale": [50000, 50000, 1]},
             "constraints": {"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9]},
                             "likelihood_variance": {"low": 0.00125, "high": 0.01}}}
    pred = {"method": "from_dataframe", "df": ploc, "max_dist": 250_000}
    oi = LocalExpertOI(expert_loc_config={"source": eloc}, data_config=data, model_config=model, pred_loc_config=pred)
    tabs = oi.run(store_path=None, optimise=True, predict=False)
    ref_model = {k: v for k, v in model.items() if k != "oi_model"}
    ref_tabs, _ = run_local_expert_oi(eloc, data, ref_model, pred, optimise=True)
    f, fr = tabs["run_details"]["objective_value"].values, ref_tabs["run_details"]["objective_value"].values
    assert (f <= fr + 1e-6 * np.abs(fr)).all()
    # --- smooth the GPU run's hyper-parameter tables on the GPU; the oracle smooths the same tables
    smoothed, ref_smoothed = {}, {}
    cfg = {"lengthscales": dict(l_x=200_000, l_y=200_000, max=12), "kernel_variance": dict(l_x=200_000, l_y=200_000),
           "likelihood_variance": dict(l_x=200_000, l_y=200_000, min=1e-4)}
    for nm, c in cfg.items():
        df = tabs[nm].reset_index()
        smoothed[f"{nm}_SMOOTHED"] = pp.smooth_hyperparameter_table(df, nm, ["x", "y", "t"], **c)
        ref_smoothed[nm] = opp.smooth_table(df, nm, ["x", "y", "t"], **c)
        assert smoothed[f"{nm}_SMOOTHED"].index.equals(ref_smoothed[nm].index)
        np.testing.assert_allclose(smoothed[f"{nm}_SMOOTHED"][nm].values, ref_smoothed[nm][nm].values, rtol=1e-11)
    # --- predict-only run with the smoothed parameters (config-2 shape)
    oi2 = LocalExpertOI(expert_loc_config={"source": eloc}, data_config=data,
                        model_config=dict(model, load_params={"file": smoothed, "table_suffix": "_SMOOTHED"}),
                        pred_loc_config=pred)
    t2 = oi2.run(store_path=None, optimise=False, table_suffix="_SMOOTHED")

    def lp(row):
        out = {}
        for nm in cfg:
            d = ref_smoothed[nm].reset_index()
            sel = d[(d["x"] == row["x"]) & (d["y"] == row["y"]) & (d["t"] == row["t"])]
            out[nm] = sel.sort_values("_dim_0")[nm].values if nm == "lengthscales" else float(sel[nm].values[0])
        return out

    r2, _ = run_local_expert_oi(eloc, data, ref_model, pred, optimise=False, load_params=lp)
    p, pr = t2["preds_SMOOTHED"], r2["preds"]
    assert p.index.equals(pr.index)
    for c in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(p[c].values, pr[c].values, rtol=1e-8, atol=1e-12)
    # --- glue the overlapping expert predictions into one field
    pf, prf = p.reset_index(), pr.reset_index()
    glued = pp.glue_local_predictions_2d(pf, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"], 250_000.0)
    ref_glued = opp.glue_local_predictions(prf, ["pred_loc_x", "pred_loc_y"], ["x", "y"], ["f*", "f*_var"], 250_000.0)
    assert len(glued) == len(ref_glued) == len(ploc)
    np.testing.assert_allclose(glued.values, ref_glued.values, rtol=1e-8, atol=1e-12)
    # the glued field reproduces the smooth signal far better than the 0.05 observation noise
    truth = 0.1 * np.sin(glued["pred_loc_x"] / 2e5) + 0.05 * np.cos(glued["pred_loc_y"] / 1.5e5)
    mean_obs = obs["z"].mean()       # obs_mean is not 'local' here: f* is the field itself
    assert np.sqrt(np.mean((glued["f*"].values - truth) ** 2)) < 0.02, mean_obs
