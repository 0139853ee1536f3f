"""GPU: the batched LocalExpertOI driver against the oracle's sequential loop (same tables), and the
B200GPRModel class used one expert at a time like the reference's loop uses GPflowGPRModel."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import gpr  # noqa: E402
from oracle.local_expert_oi import run_local_expert_oi  # noqa: E402


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpsat_b200 import build
    build.build()


def _problem(seed=3, n=6000):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"x": rng.uniform(-5e5, 5e5, n), "y": rng.uniform(-5e5, 5e5, n),
                       "t": rng.integers(18320, 18333, n).astype(float)})
    df["z"] = 0.1 * np.sin(df["x"] / 2e5) + 0.05 * np.cos(df["y"] / 1.5e5) + rng.normal(0, 0.05, n)
    df["lat"] = 60 + 30 * (1 - np.hypot(df["x"], df["y"]) / 8e5)
    df["date"] = pd.to_datetime(df["t"], unit="D")
    # experts: two dates (two global where groups), one with no data nearby, one with no prediction locations
    eloc = pd.DataFrame({"x": [0.0, 2e5, -1e5, 9e6, 1e5, 3e5], "y": [0.0, -1e5, 2e5, 9e6, 1e5, 3e5],
                         "t": [18326.0, 18326.0, 18327.0, 18326.0, 18327.0, 18326.0]})
    gx, gy = np.meshgrid(np.arange(-3e5, 2.5e5 + 1, 2.5e4), np.arange(-3e5, 2.5e5 + 1, 2.5e4))
    ploc = pd.DataFrame({"x": np.r_[gx.ravel(), 9e6], "y": np.r_[gy.ravel(), 9e6]})
    data = {"data_source": df, "obs_col": "z", "coords_col": ["x", "y", "t"],
            "local_select": [{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
                             {"col": ["x", "y"], "comp": "<", "val": 150_000}],
            "global_select": [{"col": "lat", "comp": ">=", "val": 60},
                              {"loc_col": "t", "src_col": "date",
                               "func": "lambda x,y: np.datetime64(pd.to_datetime(x+y, unit='D'))"}]}
    model = {"oi_model": "B200GPRModel", "init_params": {"coords_scale": [50000, 50000, 1]},
             "constraints": {"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9]},
                             "likelihood_variance": {"low": 0.00125, "high": 0.01}}}
    pred = {"method": "from_dataframe", "df": ploc, "max_dist": 100_000}
    return eloc, data, model, pred


def test_run_tables_match_sequential_oracle(cuda):
    from gpsat_b200.local_experts import LocalExpertOI
    eloc, data, model, pred = _problem()
    oi = LocalExpertOI(expert_loc_config={"source": eloc}, data_config=data, model_config=model,
                       pred_loc_config=pred)
    tabs = oi.run(store_path=None, optimise=True, min_obs=3)
    # oracle: sequential loop; data_source filtered by the static + dynamic where per expert inside
    ref_tabs, per = run_local_expert_oi(eloc, data, {k: v for k, v in model.items() if k != "oi_model"}, pred)
    assert set(ref_tabs) <= set(tabs)
    rd, rrd = tabs["run_details"], ref_tabs["run_details"]
    assert list(rd.index.names) == ["x", "y", "t"]
    assert rd.index.equals(rrd.index)
    np.testing.assert_array_equal(rd["num_obs"].values, rrd["num_obs"].values)
    np.testing.assert_array_equal(rd["optimise_success"].values, rrd["optimise_success"].values)
    assert list(rd.columns) == list(rrd.columns)
    ok = ~np.isnan(rrd["objective_value"].values)
    assert (np.isnan(rd["objective_value"].values) == ~ok).all()
    f, fr = rd["objective_value"].values[ok], rrd["objective_value"].values[ok]
    assert (f <= fr + 1e-6 * np.abs(fr)).all()
    assert (rd["device"].values[~ok] == "").all() and rd["model"].iloc[0] == "gpsat_b200.model.B200GPRModel"
    for nm in ("lengthscales", "kernel_variance", "likelihood_variance"):
        assert tabs[nm].index.equals(ref_tabs[nm].index)
        assert list(tabs[nm].columns) == list(ref_tabs[nm].columns)
        np.testing.assert_array_equal(tabs[nm]["_dim_0"].values, ref_tabs[nm]["_dim_0"].values)
    p, pr = tabs["preds"], ref_tabs["preds"]
    assert list(p.columns) == list(pr.columns)
    assert p.index.equals(pr.index)
    for c in ("pred_loc_x", "pred_loc_y", "pred_loc_t", "_dim_0", "f_bar"):
        np.testing.assert_array_equal(p[c].values, pr[c].values)
    for c in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(p[c].values, pr[c].values, rtol=1e-4, atol=1e-4 * np.abs(pr[c].values).max())
    assert "expert_locs" in tabs and "oi_config" in tabs


def test_predict_only_with_loaded_parameters(cuda):
    """config-2 shape: optimise=False, parameters loaded per expert, 1e-8 prediction parity."""
    from gpsat_b200.local_experts import LocalExpertOI
    eloc, data, model, pred = _problem(seed=5)
    eloc = eloc.iloc[[0, 1, 5]].reset_index(drop=True)
    rng = np.random.default_rng(0)
    E = len(eloc)
    ls = rng.uniform(2, 8, (E, 3))
    kv = rng.uniform(0.005, 0.03, E)
    nv = rng.uniform(0.002, 0.008, E)
    idx = pd.MultiIndex.from_arrays([eloc["x"], eloc["y"], eloc["t"]], names=["x", "y", "t"])
    src = {"lengthscales_SMOOTHED": pd.DataFrame({"_dim_0": np.tile(np.arange(3), E), "lengthscales": ls.ravel()},
                                                 index=idx.repeat(3)),
           "kernel_variance_SMOOTHED": pd.DataFrame({"_dim_0": 0, "kernel_variance": kv}, index=idx),
           "likelihood_variance_SMOOTHED": pd.DataFrame({"_dim_0": 0, "likelihood_variance": nv}, index=idx)}
    model = dict(model, load_params={"file": src, "table_suffix": "_SMOOTHED"})
    oi = LocalExpertOI(expert_loc_config={"source": eloc}, data_config=data, model_config=model,
                       pred_loc_config=pred)
    tabs = oi.run(store_path=None, optimise=False, table_suffix="_SMOOTHED")

    def lp(row):
        k = int(np.flatnonzero((eloc["x"] == row["x"]) & (eloc["y"] == row["y"]) & (eloc["t"] == row["t"]))[0])
        return {"lengthscales": ls[k], "kernel_variance": kv[k], "likelihood_variance": nv[k]}

    ref_tabs, per = run_local_expert_oi(eloc, data, {k: v for k, v in model.items()
                                                     if k not in ("oi_model", "load_params")}, pred,
                                        optimise=False, load_params=lp)
    rd, rrd = tabs["run_details_SMOOTHED"], ref_tabs["run_details"]
    np.testing.assert_allclose(rd["objective_value"].values, rrd["objective_value"].values, rtol=1e-8)
    assert not rd["optimise_success"].any() and not rd["parameters_optimised"].any()
    p, pr = tabs["preds_SMOOTHED"], ref_tabs["preds"]
    assert p.index.equals(pr.index)
    for c in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(p[c].values, pr[c].values, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(tabs["lengthscales_SMOOTHED"]["lengthscales"].values,
                               ref_tabs["lengthscales"]["lengthscales"].values, rtol=1e-14)


def test_model_class_like_the_reference_loop(cuda):
    """One model per expert, exactly the calls LocalExpertOI.run makes (local_experts.py:1043-1159)."""
    from gpsat_b200.model import B200GPRModel
    rng = np.random.default_rng(7)
    n = 180
    df = pd.DataFrame({"x": rng.uniform(-3e5, 3e5, n), "y": rng.uniform(-3e5, 3e5, n),
                       "t": rng.integers(18322, 18331, n).astype(float)})
    df["obs"] = 0.1 * np.sin(df["x"] / 2e5) + rng.normal(0, 0.05, n)
    cons = {"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9], "scale": True},
            "likelihood_variance": {"low": 0.00125, "high": 0.01}}
    kw = dict(data=df, obs_col="obs", coords_col=["x", "y", "t"], coords_scale=[50000, 50000, 1], obs_mean="local")
    m = B200GPRModel(expert_loc=np.zeros(3), verbose=False, **kw)
    o = gpr.OracleGPRModel(**kw)
    for mod in (m, o):
        mod.set_parameter_constraints({k: dict(v) for k, v in cons.items()}, move_within_tol=True, tol=1e-2)
    f0, f0r = m.get_objective_function_value(), o.get_objective_function_value()
    assert abs(f0 - f0r) <= 1e-8 * abs(f0r)
    Xp = np.column_stack([rng.uniform(-2e5, 2e5, (30, 2)), np.full(30, 18326.0)])
    p0, p0r = m.predict(coords=Xp), o.predict(Xp)
    for k in ("f*", "f*_var", "y_var", "f_bar"):
        np.testing.assert_allclose(p0[k], p0r[k], rtol=1e-8, atol=1e-12)
    ok, okr = m.optimise_parameters(), o.optimise_parameters()
    assert ok == okr
    f1, f1r = m.get_objective_function_value(), o.get_objective_function_value()
    assert f1 <= f1r + 1e-6 * abs(f1r)
    p1, p1r = m.predict(coords=pd.DataFrame(Xp, columns=["x", "y", "t"])), o.predict(Xp)
    for k in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(p1[k], p1r[k], rtol=1e-4, atol=1e-4 * np.abs(p1r[k]).max())
    hy, hyr = m.get_parameters(), o.get_parameters()
    np.testing.assert_allclose(hy["lengthscales"], hyr["lengthscales"], rtol=1e-3)
    # fixed parameters stay fixed
    m2 = B200GPRModel(verbose=False, **kw)
    m2.set_parameters(likelihood_variance=0.004)
    m2.optimise_parameters(fixed_params=["likelihood_variance"])
    assert m2.get_likelihood_variance() == 0.004


def test_model_full_cov(cuda):
    from gpsat_b200.model import B200GPRModel
    rng = np.random.default_rng(8)
    n = 150
    X = np.column_stack([rng.uniform(-3e5, 3e5, (n, 2)), rng.integers(18322, 18331, n).astype(float)])
    z = 0.1 * np.sin(X[:, 0] / 2e5) + rng.normal(0, 0.05, n)
    kw = dict(coords=X, obs=z, coords_scale=[50000, 50000, 1], obs_mean="local",
              kernel_kwargs={"lengthscales": [3.0, 4.0, 5.0], "variance": 0.02}, noise_variance=0.004)
    m, o = B200GPRModel(verbose=False, **kw), gpr.OracleGPRModel(**kw)
    for P in (7, 64, 150):
        Xp = np.column_stack([rng.uniform(-2e5, 2e5, (P, 2)), np.full(P, 18326.0)])
        pc, pcr = m.predict(coords=Xp, full_cov=True), o.predict(Xp, full_cov=True)
        for k in ("f*", "f*_var", "y_var", "f*_cov", "y_cov"):
            np.testing.assert_allclose(pc[k], pcr[k], rtol=1e-8, atol=1e-10 * np.abs(pcr[k]).max())
