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
