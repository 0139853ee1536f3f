"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the authoring container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

The reference's GPflow path cannot be imported (gpflow/tensorflow/tables/xarray/pyproj
absent), so four stub modules are installed (SURVEY.md section 8c "stub recipe") and the
pieces of the reference that DO run are executed unmodified:

  * GPSat.dataloader.DataLoader.local_data_select  (real scipy KDTree)        -> select_*.npz
  * GPSat.prediction_locations.PredictionLocations (numba _max_dist_bool)     -> predloc_*.npz
  * GPSat.models.pure_python_gpr.PurePythonGPR  (Matern-3/2 ARD, numpy)       -> gpr3d_*.npz
  * sklearn GaussianProcessRegressor exactly as tests/test_localexperts.py:22-49 (KAT-1)
    and docs/notebooks/gp_regression.ipynb (KAT-3)                            -> kat1.npz, kat3.npz
  * the 1-D selection counts of docs/notebooks/1d_local_expert_model_part_2.ipynb (KAT-4)
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tf = mod("tensorflow")
    tfp_ = mod("tensorflow.python")
    tfc = mod("tensorflow.python.client")
    dl = mod("tensorflow.python.client.device_lib", list_local_devices=lambda: [])
    tf.python, tfp_.client, tfc.device_lib = tfp_, tfc, dl
    mod("tables")

    class _DA:  # xarray.DataArray / Dataset placeholders (isinstance checks only)
        pass

    class _DS:
        pass

    xr = mod("xarray", DataArray=_DA, Dataset=_DS)
    xc = mod("xarray.core")
    xd = mod("xarray.core.dataarray", DataArray=_DA, Dataset=_DS)
    xr.core, xc.dataarray = xc, xd
    mod("pyproj", Transformer=object)
    sys.path.insert(0, REF)


def gen_selection(rng):
    from GPSat.dataloader import DataLoader
    from GPSat.prediction_locations import PredictionLocations
    # observations snapped to a 50 km lattice (exact ties on the radius) + jittered points
    n = 30000
    gx = rng.integers(-40, 41, n) * 50_000.0
    gy = rng.integers(-40, 41, n) * 50_000.0
    jit = rng.random(n) < 0.5
    gx = np.where(jit, gx + rng.normal(0, 20_000, n), gx)
    gy = np.where(jit, gy + rng.normal(0, 20_000, n), gy)
    t = rng.integers(18316, 18337, n).astype(np.float64)
    df = pd.DataFrame({"x": gx, "y": gy, "t": t, "z": rng.normal(size=n)})
    local_select = [{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
                    {"col": ["x", "y"], "comp": "<", "val": 300_000}]
    E = 40
    ex = rng.integers(-30, 31, E) * 50_000.0
    ey = rng.integers(-30, 31, E) * 50_000.0
    et = rng.integers(18320, 18333, E).astype(np.float64)
    offsets = [0]
    idx = []
    for i in range(E):
        rl = pd.DataFrame({"x": [ex[i]], "y": [ey[i]], "t": [et[i]]})
        out = DataLoader.local_data_select(df, reference_location=rl, local_select=local_select,
                                           verbose=False)
        idx.append(out.index.values.astype(np.int64))
        offsets.append(offsets[-1] + len(out))
    np.savez_compressed(os.path.join(HERE, "select_3d.npz"), x=gx, y=gy, t=t, ex=ex, ey=ey, et=et,
                        radius=300_000.0, t_lo=-4.0, t_hi=4.0,
                        offsets=np.array(offsets), idx=np.concatenate(idx))
    print("select_3d: counts", np.diff(offsets)[:10], "...")

    # prediction-location filter
    px, py = np.meshgrid(np.arange(-2_000_000, 2_000_001, 25_000.0),
                         np.arange(-2_000_000, 2_000_001, 25_000.0))
    ploc = pd.DataFrame({"x": px.ravel(), "y": py.ravel()})
    pl = PredictionLocations(method="from_dataframe", coords_col=["x", "y", "t"], df=ploc,
                             max_dist=200_000)
    poffsets = [0]
    pidx = []
    pc_first = None
    for i in range(E):
        pl.expert_loc = pd.DataFrame({"x": [ex[i]], "y": [ey[i]], "t": [et[i]]})
        pc = pl()
        # recover row indices of the kept locations
        key = {(a, b): k for k, (a, b) in enumerate(zip(ploc["x"].values, ploc["y"].values))}
        ids = np.array([key[(a, b)] for a, b in zip(pc[:, 0], pc[:, 1])], dtype=np.int64)
        assert np.all(pc[:, 2] == et[i])
        pidx.append(ids)
        poffsets.append(poffsets[-1] + len(ids))
        if pc_first is None:
            pc_first = pc
    np.savez_compressed(os.path.join(HERE, "predloc_2d.npz"), px=ploc["x"].values, py=ploc["y"].values,
                        ex=ex, ey=ey, et=et, max_dist=200_000.0,
                        offsets=np.array(poffsets), idx=np.concatenate(pidx), first=pc_first)
    print("predloc_2d: counts", np.diff(poffsets)[:10], "...")


def gen_kat4():
    """docs/notebooks/1d_local_expert_model_part_2.ipynb: selection counts on seed-0 1-D data."""
    from GPSat.dataloader import DataLoader
    # the notebook's data generation (cell 3): N=100 uniform x in [0,1], seed 0
    # notebook cell 3 verbatim: X ~ U(0.1, 0.6), seed 0, N=100
    np.random.seed(0)
    N = 100
    x = np.random.uniform(0.1, 0.6, (N,))
    y = np.sin(1 / x) + 0.05 * np.random.randn(N)
    df = pd.DataFrame({"x": x})
    out = {}
    for radius, centers in [(0.15, [0.25, 0.45]), (0.1, [0.2, 0.3, 0.4, 0.5])]:
        ls = [{"col": "x", "comp": "<=", "val": radius}, {"col": "x", "comp": ">=", "val": -radius}]
        cnt = []
        for c in centers:
            o = DataLoader.local_data_select(df, reference_location=pd.DataFrame({"x": [c]}),
                                             local_select=ls, verbose=False)
            cnt.append(len(o))
        out[str(radius)] = cnt
    print("KAT-4 counts (notebook prints 62, 59 and 41, 37, 44, 38):", out)
    assert out["0.15"] == [62, 59] and out["0.1"] == [41, 37, 44, 38]
    np.savez_compressed(os.path.join(HERE, "kat4.npz"), x=x, y=y,
                        c015=np.array([0.25, 0.45]), n015=np.array(out["0.15"]),
                        c01=np.array([0.2, 0.3, 0.4, 0.5]), n01=np.array(out["0.1"]))
    return out


def gen_gpr3d(rng):
    from GPSat.models.pure_python_gpr import PurePythonGPR
    cases = {}
    for tag, (N, P, ls, kv, nv) in {
        "a": (400, 50, [5.18430274, 3.21994817, 8.99996751], 0.015248077637888286, 0.003326551981572017),
        "b": (257, 33, [2.0, 1.5, 4.0], 0.05, 0.01),
        "c": (64, 7, [1.0, 1.0, 1.0], 1.0, 0.005625),
    }.items():
        xy = rng.integers(-6, 7, (N, 2)) * 50_000.0 + rng.normal(0, 5_000, (N, 2))
        t = rng.integers(18322, 18331, N).astype(np.float64)
        X = np.column_stack([xy, t])
        z = 0.1 * np.sin(X[:, 0] / 2e5) + 0.05 * np.cos(X[:, 1] / 1.5e5) + rng.normal(0, 0.05, N)
        Xs = np.column_stack([rng.uniform(-3e5, 3e5, (P, 2)), np.full(P, 18326.0)])
        m = PurePythonGPR(coords=X.copy(), obs=z.copy(), coords_scale=[50_000, 50_000, 1],
                          obs_mean="local", length_scales=np.array(ls), kernel_var=kv,
                          likeli_var=nv, verbose=False)
        nlml = float(np.squeeze(m.get_objective_function_value()))
        pred = m.predict(Xs)
        cases[tag] = dict(X=X, z=z, Xs=Xs, ls=np.array(ls), kv=kv, nv=nv, nlml=nlml,
                          fstar=np.asarray(pred["f*"]).ravel(), fvar=np.asarray(pred["f*_var"]).ravel())
        print(f"gpr3d_{tag}: -LML={nlml:.12f}  f*[0]={cases[tag]['fstar'][0]:.10f}  var[0]={cases[tag]['fvar'][0]:.10e}")
    np.savez_compressed(os.path.join(HERE, "gpr3d.npz"),
                        **{f"{t}_{k}": v for t, c in cases.items() for k, v in c.items()})


def gen_kat1_kat3():
    from sklearn.gaussian_process.kernels import Matern, RBF, ConstantKernel, WhiteKernel
    from sklearn.gaussian_process import GaussianProcessRegressor
    # --- KAT-1: tests/test_localexperts.py:22-49 verbatim ---
    np.random.seed(23435)
    kernel = Matern(length_scale=0.8, nu=3 / 2)
    gp = GaussianProcessRegressor(kernel)
    x = np.linspace(0, 10, 100)[:, None]
    f = gp.sample_y(x, random_state=0)
    N = 50
    eps = 1e-2
    indices = np.arange(100)
    np.random.shuffle(indices)
    x_train = x[indices[:N]]
    y_train = f[indices[:N]] + eps * np.random.randn(N, 1)
    gp.alpha = eps ** 2
    gp.fit(x_train, y_train)
    ls = gp.kernel_.length_scale
    ml = gp.log_marginal_likelihood()
    test_index = np.random.randint(0, 99)
    x_test = x[[test_index]]
    pred_mean, pred_std = gp.predict(x_test, return_std=True)
    np.savez_compressed(os.path.join(HERE, "kat1.npz"), x_train=x_train, y_train=y_train, eps=eps,
                        ls=ls, ml=ml, x_test=x_test, pred_mean=pred_mean, pred_var=pred_std ** 2)
    print(f"KAT-1: ls={ls!r} ml={ml!r} mean={pred_mean} var={pred_std**2}")
    # --- KAT-3: docs/notebooks/gp_regression.ipynb: N=30, y=cos(x)+0.05 eps, RBF l=1, amp sqrt(1.5)
    # notebook cell 3 verbatim
    np.random.seed(0)
    Nk = 30
    xk = np.random.uniform(-5, 5, (Nk,))
    yk = np.cos(xk) + 0.05 * np.random.randn(Nk)
    # sklearn_models.py:95-96 multiplies by ConstantKernel(sqrt(kernel_variance)): effective variance sqrt(1.5)
    k3 = ConstantKernel(np.sqrt(1.5), constant_value_bounds="fixed") * RBF(1.0, length_scale_bounds="fixed")
    gp3 = GaussianProcessRegressor(k3, alpha=0.0025, optimizer=None).fit(xk[:, None], yk)
    np.savez_compressed(os.path.join(HERE, "kat3.npz"), x=xk, y=yk, ls=1.0, kv=np.sqrt(1.5), nv=0.0025,
                        ml=gp3.log_marginal_likelihood(),
                        mean=gp3.predict(np.array([[0.3], [1.7]])), xs=np.array([[0.3], [1.7]]),
                        var=gp3.predict(np.array([[0.3], [1.7]]), return_std=True)[1] ** 2)
    print("KAT-3 (notebook prints 16.6180): ml =", gp3.log_marginal_likelihood())
    assert abs(gp3.log_marginal_likelihood() - 16.6180) < 5e-5


if __name__ == "__main__":
    install_stubs()
    rng = np.random.default_rng(20200305)
    gen_selection(rng)
    gen_kat4()
    gen_gpr3d(rng)
    gen_kat1_kat3()
