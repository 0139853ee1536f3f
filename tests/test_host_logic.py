"""CPU tests of the host-side mirror of the reference interface (no GPU, no oracle on the product path:
the oracle appears here only as the checker)."""
import numpy as np
import pandas as pd
import pytest

from gpsat_b200.batched import ModelSpec, sel_terms
from gpsat_b200.dataloader import data_select, get_where_list
from gpsat_b200.local_experts import pretty_print_class
from gpsat_b200.model import B200GPRModel, get_model
from gpsat_b200.params import HyperParams
from oracle import gpr


def test_constraints_match_oracle_worked_example():
    """SURVEY 8a row M4 worked example: inline config -> l in [2e-13, 12] x2, [1e-8, 9]; noise start 0.005625."""
    X = np.random.default_rng(1).normal(size=(10, 3))
    cons = {"lengthscales": {"low": [1e-8] * 3, "high": [600000, 600000, 9], "scale": True},
            "likelihood_variance": {"low": 0.00125, "high": 0.01}}
    o = gpr.OracleGPRModel(coords=X.copy(), obs=X[:, 0].copy(), coords_scale=[50_000, 50_000, 1])
    o.set_parameter_constraints({k: dict(v) for k, v in cons.items()}, move_within_tol=True, tol=1e-2)
    m = B200GPRModel(coords=X.copy(), obs=X[:, 0].copy(), coords_scale=[50_000, 50_000, 1], verbose=False)
    m.set_parameter_constraints({k: dict(v) for k, v in cons.items()}, move_within_tol=True, tol=1e-2)
    np.testing.assert_array_equal(m.get_lengthscales(), o.get_lengthscales())
    assert m.get_likelihood_variance() == o.get_likelihood_variance() == pytest.approx(0.005625, abs=1e-15)
    k1, l1, h1 = m._hp.transforms()
    k2, l2, h2 = o._transforms_flat()
    np.testing.assert_array_equal(k1, k2)
    np.testing.assert_array_equal(l1, l2)
    np.testing.assert_array_equal(h1, h2)
    # the batched spec produces the same start point and bijectors from the model config
    spec = ModelSpec.from_model_config({"init_params": {"coords_scale": [50_000, 50_000, 1]},
                                        "constraints": {k: {kk: vv for kk, vv in v.items() if kk != "scale"}
                                                        for k, v in cons.items()}})
    hp = spec.hyper_params(3)
    np.testing.assert_array_equal(hp.theta(), m._hp.theta())
    np.testing.assert_array_equal(hp.transforms()[2], h1)


def test_model_interface_without_gpu():
    """postprocessing.smooth_hyperparameters builds a model on a 1-row frame just to read param_names."""
    df = pd.DataFrame({"x": [0.0], "y": [1.0], "t": [2.0], "obs": [0.5]})
    m = B200GPRModel(data=df, coords_col=["x", "y", "t"], obs_col="obs", expert_loc=np.zeros(3), verbose=False)
    assert m.param_names == ["lengthscales", "kernel_variance", "likelihood_variance"]
    p = m.get_parameters()
    np.testing.assert_array_equal(p["lengthscales"], np.ones(3))
    assert p["kernel_variance"] == 1.0 and p["likelihood_variance"] == 1.0
    m.set_parameters(lengthscales=[2.0, 3.0, 4.0], kernel_variance=np.array([0.5]), likelihood_variance=1e-9)
    assert m.get_likelihood_variance() == 1e-6      # clipped to gpflow's lower bound
    assert m.get_parameters("kernel_variance", return_dict=False) == [0.5]
    with pytest.raises(AssertionError):
        m.get_parameters("nope")
    with pytest.raises(AssertionError):
        m.set_parameters(nope=1)
    with pytest.raises(AssertionError):
        B200GPRModel(coords=np.array([[np.nan]]), obs=np.array([1.0]))
    assert get_model("B200GPRModel") is B200GPRModel
    with pytest.raises(NotImplementedError):
        get_model("sklearnGPRModel")
    # obs_mean: only 'local' is honoured (base_model.py:195-200)
    m2 = B200GPRModel(coords=np.arange(4.0), obs=np.array([1.0, 2, 3, 4]), obs_mean=7.0, verbose=False)
    assert m2.obs_mean[0, 0] == 0 and m2.obs[:, 0].tolist() == [1, 2, 3, 4]
    m3 = B200GPRModel(coords=np.arange(4.0), obs=np.array([1.0, 2, 3, 4]), obs_mean="local", obs_scale=2, verbose=False)
    assert m3.obs_mean[0, 0] == 2.5 and m3.obs[:, 0].tolist() == [-0.75, -0.25, 0.25, 0.75]


def test_numerical_methods_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = B200GPRModel(coords=np.arange(4.0), obs=np.array([1.0, 2, 3, 4]), verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_objective_function_value()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.optimise_parameters()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.predict(np.array([[1.0]]))


def test_hyperparams_fixed_mask_and_roundtrip():
    hp = HyperParams(3, [1, 2, 3], 0.5, 0.1)
    np.testing.assert_array_equal(hp.trainable_mask(["likelihood_variance", "kernel_variance"]), [1, 1, 1, 0, 0])
    np.testing.assert_array_equal(hp.trainable_mask(["lengthscales"]), [0, 0, 0, 1, 1])
    th = hp.theta()
    hp2 = HyperParams(3)
    hp2.set_theta(th)
    np.testing.assert_array_equal(hp2.theta(), th)
    k, lo, hi = hp.transforms()
    assert k.tolist() == [0] * 5 and lo[4] == 1e-6
    with pytest.raises(AssertionError):
        hp.set_constraints("lengthscales", [1, 1], [2, 2])


def test_sel_terms_and_where_list():
    ls = [{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
          {"col": ["x", "y"], "comp": "<", "val": 300_000}]
    t = sel_terms(ls, ["x", "y", "t", "obs"], ["x", "y", "t"])
    assert [q["type"] for q in t] == [0, 0, 1]
    assert t[0]["cols"] == [2] and t[2]["cols"] == [0, 1] and t[2]["val"] == 300_000
    gs = [{"col": "lat", "comp": ">=", "val": 60},
          {"loc_col": "t", "src_col": "date", "func": "lambda x,y: np.datetime64(pd.to_datetime(x+y, unit='D'))"}]
    w = get_where_list(gs, ls, {"x": 0.0, "y": 0.0, "t": 18326.0})
    assert w[0] == gs[0]
    assert w[1] == {"col": "date", "comp": "<=", "val": np.datetime64("2020-03-09")}
    assert w[2] == {"col": "date", "comp": ">=", "val": np.datetime64("2020-03-01")}
    df = pd.DataFrame({"lat": [50.0, 70.0, 80.0],
                       "date": pd.to_datetime(["2020-03-05", "2020-03-05", "2020-04-01"])})
    assert data_select(df, where=w).index.tolist() == [1]
    assert pretty_print_class(B200GPRModel) == "gpsat_b200.model.B200GPRModel"


def test_register_wraps_get_model_in_all_three_namespaces():
    """GPSat.models.get_model is imported by value into GPSat.local_experts and GPSat.postprocessing
    (models/__init__.py:3, local_experts.py:32, postprocessing.py:18): register() must wrap all three."""
    import sys
    import types
    from gpsat_b200 import model as bm

    def ref_get_model(name):
        if name == "sklearnGPRModel":
            return "sk"
        raise NotImplementedError(f"model with name: '{name}' is not implemented")

    saved = {k: sys.modules.get(k) for k in ("GPSat", "GPSat.models", "GPSat.local_experts", "GPSat.postprocessing")}
    try:
        pkg = types.ModuleType("GPSat")
        sys.modules["GPSat"] = pkg
        mods = []
        for nm in ("models", "local_experts", "postprocessing"):
            m = types.ModuleType(f"GPSat.{nm}")
            m.get_model = ref_get_model
            sys.modules[f"GPSat.{nm}"] = m
            setattr(pkg, nm, m)
            mods.append(m)
        bm.register()
        for m in mods:
            assert m.get_model("B200GPRModel") is bm.B200GPRModel
            assert m.get_model("B200SGPRModel") is bm.B200SGPRModel
            assert m.get_model("sklearnGPRModel") == "sk"          # everything else still goes to the reference
            with pytest.raises(NotImplementedError):
                m.get_model("nope")
        bm.register()                                               # idempotent
        assert mods[0].get_model("B200GPRModel") is bm.B200GPRModel
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_result_buffers_are_recycled_only_when_unreferenced():
    """batched.to_host hands out views of one pageable buffer per call and reuses a buffer for the next call only when
    nothing references it any more (views of views included); a caller that keeps any piece gets a fresh buffer."""
    import numpy as np
    from gpsat_b200 import batched as b
    b._HOST_POOL.clear()
    buf = b._host_buffer(4096)
    a = buf[:800].view(np.float64)
    piece = a[10:20]                      # a view of a view still pins the owner
    a[:] = 7.0
    del buf, a
    other = b._host_buffer(1024)
    assert not np.shares_memory(other, piece)            # still referenced: not handed out again
    assert piece[0] == 7.0
    del piece, other
    again = b._host_buffer(2048)
    assert any(again is x for x in b._HOST_POOL) and len(b._HOST_POOL) == 2      # one of the two idle buffers, reused
    held = [again]
    for _ in range(8):                    # a caller that keeps everything: the pool stays bounded
        held.append(b._host_buffer(1 << 20))
    assert len(b._HOST_POOL) <= b._HOST_POOL_MAX
    held[1][:] = 3                        # forgotten by the pool, alive as long as the caller holds it
    assert held[1][0] == 3
    b._HOST_POOL.clear()


def test_pinned_result_buffers_alternate_and_fall_back(monkeypatch):
    """batched._pinned_buffer (the zero-copy path of to_host): at most two page-locked buffers, handed out only when
    unreferenced; with both still held the caller is told to use the staging path (None)."""
    import numpy as np
    import torch
    from gpsat_b200 import batched as b
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items()
                                                                          if x != "pin_memory"}))   # no CUDA here
    b._PINNED_POOL.clear()
    t1, r1 = b._pinned_buffer(1000)
    first = r1[:80].view(np.float64)
    del t1, r1
    t2, r2 = b._pinned_buffer(1000)
    assert not np.shares_memory(first, r2) and len(b._PINNED_POOL) == 2
    second = r2[:8]
    del t2, r2
    assert b._pinned_buffer(100) is None                      # both referenced: fall back
    del first
    t3, r3 = b._pinned_buffer(500)
    assert r3 is b._PINNED_POOL[0][1]                          # the released one is reused
    del t3, r3
    t4, r4 = b._pinned_buffer(100_000)                         # too small: the idle buffer is replaced, not added
    assert len(b._PINNED_POOL) == 2 and r4.nbytes >= 100_000
    del second, t4, r4
    b._PINNED_POOL.clear()


def test_shape_tables_prediction_block_matches_per_expert_loop():
    """LocalExpertOI._shape_tables gathers the prediction columns of a whole batch into one block (slice when the kept
    experts' rows are consecutive, index map otherwise) and builds the MultiIndex from per-expert codes: same table as
    the reference's per-expert dict_of_array_to_table loop (local_experts.py:691-747), for both row layouts and both
    f_bar forms (float for obs_mean='local', int64 zeros otherwise, base_model.py:199-200)."""
    import numpy as np
    import pandas as pd
    from gpsat_b200.local_experts import LocalExpertOI
    rng = np.random.default_rng(3)
    E, D = 7, 3
    eloc = pd.DataFrame({"x": rng.integers(-3, 3, E) * 1e5, "y": rng.integers(-3, 3, E) * 1e5, "t": 18326.0})
    cnt = np.array([4, 0, 6, 1, 3, 5, 2])
    for local in (True, False):
        for valid_idx in (np.arange(E), np.array([2, 0, 1, 6, 3, 5, 4])):        # consecutive rows / permuted rows
            poff = np.r_[0, np.cumsum(cnt[valid_idx])]       # CSR over the VALID list, in valid_idx order
            n = int(poff[-1])
            res = dict(num_obs=rng.integers(5, 9, E), too_few=np.zeros(E, bool), valid_idx=valid_idx,
                       fobj=rng.normal(size=E), status=np.ones(E, int), theta=rng.uniform(1, 2, (E, D + 2)),
                       pred_offsets=poff, obs_mean=rng.normal(size=E), fmean=rng.normal(size=n),
                       fvar=rng.uniform(size=n), yvar=rng.uniform(size=n), pred_coords=rng.normal(size=(n, D)))
            oi = LocalExpertOI.__new__(LocalExpertOI)
            oi.coords_col = ["x", "y", "t"]
            oi.params_to_store = None
            oi.model_config = {"init_params": {"obs_mean": "local"} if local else {}}
            pieces = {}
            oi._shape_tables(pieces, res, None, eloc, np.arange(E), np.ones(E, bool), 0.1, True, True, "m", "d", 1, D,
                             True)
            got = pieces["preds"][0]
            rows = []
            for e in range(E):                              # global expert order, one frame per expert
                v = int(np.flatnonzero(valid_idx == e)[0])
                sl = slice(poff[v], poff[v + 1])
                c = poff[v + 1] - poff[v]
                fb = np.full(c, res["obs_mean"][v]) if local else np.zeros(c, dtype=np.int64)
                d = pd.DataFrame({"_dim_0": np.arange(c), "f*": res["fmean"][sl], "f*_var": res["fvar"][sl],
                                  "y_var": res["yvar"][sl], "f_bar": fb,
                                  "pred_loc_x": res["pred_coords"][sl, 0], "pred_loc_y": res["pred_coords"][sl, 1],
                                  "pred_loc_t": res["pred_coords"][sl, 2], "_pos_": e},
                                 index=pd.MultiIndex.from_arrays([np.full(c, eloc[k][e]) for k in "xyt"],
                                                                 names=list("xyt")))
                rows.append(d)
            want = pd.concat(rows, axis=0)
            pd.testing.assert_frame_equal(got, want, check_dtype=True)
            assert got.index.equals(want.index) and list(got.index.names) == ["x", "y", "t"]
