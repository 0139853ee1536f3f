"""GPU parity of the batched sparse GPR (row SG1) against the oracle restating gpflow.models.SGPR."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import sgpr  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpsat_b200 import build, get_engine
    build.build()
    return get_engine(0)


def _synth(rng, n):
    xy = rng.integers(-6, 7, (n, 2)) * 50_000.0 + rng.normal(0, 5_000, (n, 2))
    t = rng.integers(18322, 18331, n).astype(np.float64)
    X = np.column_stack([xy, t])
    z = 0.1 * np.sin(X[:, 0] / 2e5) + 0.05 * np.cos(X[:, 1] / 1.5e5) + rng.normal(0, 0.05, n)
    return X, z, np.array([50_000.0, 50_000.0, 1.0])


def _batch(eng, rng, sizes, kernel="Matern32", obs_mean_local=True):
    Xs, ys, Zs = [], [], []
    for n, m in sizes:
        X, y, cs = _synth(rng, n)
        Xs.append(X)
        ys.append(y)
        Zs.append(X[rng.permutation(n)[:m]].copy())
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([n for n, _ in sizes])
    zoff = np.zeros(len(sizes) + 1, dtype=np.int64)
    zoff[1:] = np.cumsum([m for _, m in sizes])
    b = eng.make_batch(off, np.concatenate(Xs), np.concatenate(ys), kernel=kernel, coords_scale=cs,
                       obs_mean_local=obs_mean_local)
    return eng.make_sgpr_batch(b, zoff, np.concatenate(Zs)), Xs, ys, Zs, cs


@pytest.mark.parametrize("kernel", ["Matern32", "RBF"])
def test_sgpr_elbo_and_gradient(eng, kernel):
    rng = np.random.default_rng(2)
    sizes = [(300, 64), (150, 40), (500, 130), (90, 90), (257, 65), (64, 1)]
    sb, Xs, ys, Zs, cs = _batch(eng, rng, sizes, kernel)
    E = len(sizes)
    theta = np.column_stack([rng.uniform(1.0, 6.0, E), rng.uniform(1.0, 6.0, E), rng.uniform(2.0, 9.0, E),
                             rng.uniform(0.01, 0.1, E), rng.uniform(0.002, 0.01, E)])
    f, g = eng.sgpr_eval(sb, theta, grad=True)
    f, g = f.cpu().numpy(), g.cpu().numpy()
    f2, _ = eng.sgpr_eval(sb, theta, grad=False)
    for e in range(E):
        y = ys[e] - ys[e].mean()
        fr, gr = sgpr.neg_elbo_and_grad(Xs[e] / cs, y, Zs[e] / cs, theta[e, :3], theta[e, 3], theta[e, 4], kernel)
        assert abs(f[e] - fr) <= 1e-8 * abs(fr), (e, f[e], fr)
        assert abs(f2[e].item() - fr) <= 1e-8 * abs(fr)
        np.testing.assert_allclose(g[e], gr, rtol=2e-6, atol=2e-6 * np.abs(gr).max(), err_msg=f"expert {e}")


def test_sgpr_predict(eng):
    rng = np.random.default_rng(5)
    sizes = [(300, 64), (150, 128), (420, 100)]
    sb, Xs, ys, Zs, cs = _batch(eng, rng, sizes)
    E = len(sizes)
    theta = np.column_stack([rng.uniform(2.0, 6.0, E), rng.uniform(2.0, 6.0, E), rng.uniform(3.0, 9.0, E),
                             rng.uniform(0.01, 0.05, E), rng.uniform(0.002, 0.01, E)])
    Ps = [np.column_stack([rng.uniform(-3e5, 3e5, (p, 2)), np.full(p, 18326.0)]) for p in (70, 1, 129)]
    poff = np.zeros(E + 1, dtype=np.int64)
    poff[1:] = np.cumsum([len(p) for p in Ps])
    fm, fv, yv, fo = eng.sgpr_predict(sb, theta, poff, np.concatenate(Ps))
    fm, fv, yv, fo = fm.cpu().numpy(), fv.cpu().numpy(), yv.cpu().numpy(), fo.cpu().numpy()
    for e in range(E):
        y = ys[e] - ys[e].mean()
        m, v, yvr = sgpr.predict(Xs[e] / cs, y, Zs[e] / cs, Ps[e] / cs, theta[e, :3], theta[e, 3], theta[e, 4])
        sl = slice(poff[e], poff[e + 1])
        np.testing.assert_allclose(fm[sl], m, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(fv[sl], v, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(yv[sl], yvr, rtol=1e-6, atol=1e-9)
        fr = -sgpr.elbo(Xs[e] / cs, y, Zs[e] / cs, theta[e, :3], theta[e, 3], theta[e, 4])
        assert abs(fo[e] - fr) <= 1e-8 * abs(fr)


def test_sgpr_optimise_matches_oracle(eng):
    rng = np.random.default_rng(9)
    sizes = [(260, 60), (180, 50), (320, 100)]
    sb, Xs, ys, Zs, cs = _batch(eng, rng, sizes)
    models = []
    for e in range(len(sizes)):
        m = sgpr.OracleSGPRModel(coords=Xs[e].copy(), obs=ys[e].copy(), coords_scale=list(cs), obs_mean="local",
                                 inducing_points=Zs[e] / cs)
        m.set_parameter_constraints({"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9], "scale": True},
                                     "likelihood_variance": {"low": 0.00125, "high": 0.01}},
                                    move_within_tol=True, tol=1e-2)
        models.append(m)
    m0 = models[0]
    theta0 = np.concatenate([m0.get_lengthscales(), [m0.get_kernel_variance()], [m0.get_likelihood_variance()]])
    kind, low, high = m0._transforms_flat()
    res = eng.sgpr_optimise(sb, theta0, kind, low, high, trainable=[1] * 5)
    theta, fobj, status = res["theta"].cpu().numpy(), res["fobj"].cpu().numpy(), res["status"].cpu().numpy()
    P = 25
    Xp = np.column_stack([rng.uniform(-3e5, 3e5, (P, 2)), np.full(P, 18326.0)])
    poff = np.arange(len(sizes) + 1, dtype=np.int64) * P
    fm, fv, _, _ = eng.sgpr_predict(sb, theta, poff, np.tile(Xp, (len(sizes), 1)))
    fm, fv = fm.cpu().numpy(), fv.cpu().numpy()
    for e, m in enumerate(models):
        ok = m.optimise_parameters()
        elbo_ref = m.get_objective_function_value()
        assert -fobj[e] >= elbo_ref - 1e-6 * abs(elbo_ref), (e, -fobj[e], elbo_ref)
        assert (status[e] in (1, 2)) == ok
        out = m.predict(Xp)
        sl = slice(poff[e], poff[e + 1])
        np.testing.assert_allclose(fm[sl], out["f*"], rtol=1e-4, atol=1e-4 * np.abs(out["f*"]).max())
        np.testing.assert_allclose(fv[sl], out["f*_var"], rtol=1e-4, atol=1e-4 * np.abs(out["f*_var"]).max())


def test_kat2_sgpr_all_points_equals_sklearn(eng, golden_dir):
    """tests/test_localexperts.py:229-251: M = N = 50 -> optimised l, f*, f*_var equal sklearn's to 1e-4."""
    g = np.load(os.path.join(golden_dir, "kat1.npz"))
    x, y = g["x_train"], g["y_train"][:, 0]
    off = np.array([0, 50], dtype=np.int64)
    b = eng.make_batch(off, x, y, kernel="Matern32")
    Z = x[np.random.default_rng(0).permutation(50)]
    sb = eng.make_sgpr_batch(b, off, Z)
    nv = float(g["eps"]) ** 2
    res = eng.sgpr_optimise(sb, np.array([1.0, 1.0, nv]), kind=[1, 0, 0], low=[1e-10, 0.0, 1e-6],
                            high=[5.0, 0.0, 0.0], trainable=[1, 0, 0])
    th = res["theta"].cpu().numpy()[0]
    assert int(res["status"][0]) in (1, 2)
    assert abs(th[0] - float(g["ls"])) < 1e-4
    fm, fv, _, _ = eng.sgpr_predict(sb, th, np.array([0, 1]), g["x_test"])
    assert abs(fm.item() - float(g["pred_mean"][0])) < 1e-4
    assert abs(fv.item() - float(g["pred_var"][0])) < 1e-4


def test_sgpr_through_the_driver_matches_sequential_oracle(eng):
    """LocalExpertOI with oi_model = B200SGPRModel vs the oracle loop with OracleSGPRModel: same inducing points
    (same numpy seed and expert order), ELBO / hyper-parameters / predictions within tolerance."""
    import pandas as pd
    from gpsat_b200.local_experts import LocalExpertOI
    from oracle.local_expert_oi import run_local_expert_oi
    rng = np.random.default_rng(11)
    n = 5000
    df = pd.DataFrame({"x": rng.uniform(-5e5, 5e5, n), "y": rng.uniform(-5e5, 5e5, n),
                       "t": rng.integers(18322, 18331, n).astype(float)})
    df["z"] = 0.1 * np.sin(df["x"] / 2e5) + 0.05 * np.cos(df["y"] / 1.5e5) + rng.normal(0, 0.05, n)
    eloc = pd.DataFrame({"x": [0.0, 2e5, -1e5], "y": [0.0, -1e5, 2e5], "t": [18326.0] * 3})
    gx, gy = np.meshgrid(np.arange(-2e5, 2e5 + 1, 5e4), np.arange(-2e5, 2e5 + 1, 5e4))
    ploc = pd.DataFrame({"x": gx.ravel(), "y": gy.ravel()})
    data = {"data_source": df, "obs_col": "z", "coords_col": ["x", "y", "t"],
            "local_select": [{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
                             {"col": ["x", "y"], "comp": "<", "val": 200_000}]}
    model = {"oi_model": "B200SGPRModel",
             "init_params": {"coords_scale": [50000, 50000, 1], "num_inducing_points": 80, "obs_mean": "local"},
             "constraints": {"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9]},
                             "likelihood_variance": {"low": 0.00125, "high": 0.01}}}
    pred = {"method": "from_dataframe", "df": ploc, "max_dist": 100_000}
    np.random.seed(42)
    tabs = LocalExpertOI(expert_loc_config={"source": eloc}, data_config=data, model_config=model,
                         pred_loc_config=pred).run(store_path=None)
    np.random.seed(42)
    ref_tabs, per = run_local_expert_oi(eloc, data, {k: v for k, v in model.items() if k != "oi_model"}, pred,
                                        model_cls=sgpr.OracleSGPRModel)
    assert set(ref_tabs) <= set(tabs)
    ip, ipr = tabs["inducing_points"], ref_tabs["inducing_points"]
    assert list(ip.columns) == list(ipr.columns) and ip.index.equals(ipr.index)
    np.testing.assert_array_equal(ip["_dim_0"].values, ipr["_dim_0"].values)
    np.testing.assert_array_equal(ip["_dim_1"].values, ipr["_dim_1"].values)
    np.testing.assert_allclose(ip["inducing_points"].values, ipr["inducing_points"].values, rtol=1e-15)
    rd, rrd = tabs["run_details"], ref_tabs["run_details"]
    np.testing.assert_array_equal(rd["num_obs"].values, rrd["num_obs"].values)
    f, fr = rd["objective_value"].values, rrd["objective_value"].values       # +ELBO for the sparse model
    assert (f >= fr - 1e-6 * np.abs(fr)).all(), (f, fr)
    np.testing.assert_array_equal(rd["optimise_success"].values, rrd["optimise_success"].values)
    p, pr = tabs["preds"], ref_tabs["preds"]
    assert p.index.equals(pr.index)
    for c in ("f*", "f*_var", "y_var"):
        np.testing.assert_allclose(p[c].values, pr[c].values, rtol=1e-4, atol=1e-4 * np.abs(pr[c].values).max())
    np.testing.assert_allclose(p["f_bar"].values, pr["f_bar"].values, rtol=1e-12)


def test_sgpr_model_class(eng):
    from gpsat_b200.model import B200SGPRModel
    rng = np.random.default_rng(21)
    X, y, cs = _synth(rng, 220)
    kw = dict(coords=X, obs=y, coords_scale=list(cs), obs_mean="local", num_inducing_points=60)
    np.random.seed(5)
    m = B200SGPRModel(verbose=False, **kw)
    np.random.seed(5)
    o = sgpr.OracleSGPRModel(**kw)
    np.testing.assert_array_equal(m.get_inducing_points(), o.get_inducing_points())
    for mod in (m, o):
        mod.set_parameters(lengthscales=[3.0, 4.0, 5.0], kernel_variance=0.02, likelihood_variance=0.004)
    assert abs(m.get_objective_function_value() - o.get_objective_function_value()) <= \
        1e-8 * abs(o.get_objective_function_value())
    Xp = np.column_stack([rng.uniform(-2e5, 2e5, (30, 2)), np.full(30, 18326.0)])
    p, pr = m.predict(Xp), o.predict(Xp)
    for k in ("f*", "f*_var", "y_var", "f_bar"):
        np.testing.assert_allclose(p[k], pr[k], rtol=1e-6, atol=1e-9)
    assert m.optimise_parameters(fixed_params=["likelihood_variance"]) == o.optimise_parameters(
        fixed_params=["likelihood_variance"])
    assert m.get_likelihood_variance() == 0.004
    assert m.get_objective_function_value() >= o.get_objective_function_value() - 1e-6 * abs(
        o.get_objective_function_value())
