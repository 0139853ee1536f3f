"""B200GPRModel: the engine behind the reference's BaseGPRModel interface.

Drop-in for ``GPSat.models.gpflow_models.GPflowGPRModel`` (gpflow_models.py:26-663) as LocalExpertOI
uses it: same constructor arguments, ``param_names``, ``get_/set_parameters``,
``set_parameter_constraints``, ``optimise_parameters``, ``get_objective_function_value`` and
``predict`` (base_model.py:17-448).  It is duck-typed on purpose: the reference's BaseGPRModel
imports tensorflow at module import (base_model.py:8) and LocalExpertOI.run never isinstance-checks
the model (local_experts.py:844-846).

One instance = one expert = a batch of one on the CUDA engine.  There is no CPU fallback: every
numerical method needs the CUDA library and a device; construction and parameter bookkeeping do not
(postprocessing.smooth_hyperparameters builds a model on a 1-row dummy frame just to read
``param_names``, postprocessing.py:199-213).
"""
from __future__ import annotations

import platform
import re
import time
from typing import Dict, List, Optional, Union

import numpy as np
import pandas as pd

from .params import HyperParams, PARAM_NAMES
from ._lib import KERNEL_IDS


def _processor_name():
    try:
        if platform.system() == "Linux":
            with open("/proc/cpuinfo") as f:
                for line in f:
                    if "model name" in line:
                        return re.sub(".*model name.*:", "", line, count=1).strip()
        return platform.processor()
    except Exception:
        return "unknown"


class B200GPRModel:
    """Exact GP regression for one local expert, evaluated by hand-written sm_100a CUDA."""

    # LocalExpertOI.run may hand the whole expert list to gpsat_b200.local_experts
    supports_batched_dispatch = True

    def __init__(self,
                 data: Optional[pd.DataFrame] = None,
                 coords_col: Union[str, List[str], None] = None,
                 obs_col: Union[str, List[str], None] = None,
                 coords: Optional[np.ndarray] = None,
                 obs: Optional[np.ndarray] = None,
                 coords_scale=None, obs_scale=None, obs_mean=None, verbose: bool = True,
                 *, kernel: str = "Matern32", kernel_kwargs: Optional[dict] = None,
                 mean_function=None, mean_func_kwargs=None, noise_variance=None, likelihood=None,
                 device: int = 0, **kwargs):
        # ---- base_model.py:134-245 ----
        if data is not None:
            assert coords_col is not None, f"data was provided, but coord_col was not"
            assert obs_col is not None, f"data was provided, but obs_col was not"
            if isinstance(coords_col, str):
                coords_col = [coords_col]
            if isinstance(obs_col, str):
                obs_col = [obs_col]
            self.obs = np.array(data.loc[:, obs_col].values, dtype=np.float64)
            self.coords = np.array(data.loc[:, coords_col].values, dtype=np.float64)
        else:
            assert coords is not None and obs is not None, "either data or (coords, obs) must be given"
            self.obs = np.array(obs, dtype=np.float64)
            self.coords = np.array(coords, dtype=np.float64)
            if self.obs.ndim == 1:
                self.obs = self.obs[:, None]
            if self.coords.ndim == 1:
                self.coords = self.coords[:, None]
            assert len(self.obs) == len(self.coords), "obs and coords lengths don't match "
        self.coords_col, self.obs_col = coords_col, obs_col
        assert self.obs.shape[1] == 1, "the batched engine handles a single observation column"
        assert not np.isnan(self.coords).any(), "nans found in coords"
        assert not np.isnan(self.obs).any(), "nans found in obs"
        if isinstance(obs_mean, str) and obs_mean == "local":
            obs_mean = np.mean(self.obs, axis=0)
        else:                                    # base_model.py:199-200: everything else becomes 0
            obs_mean = np.array([0])[None, :]
        self.obs_mean = np.atleast_2d(obs_mean)

        def as2d(v):
            if v is None:
                return np.atleast_2d(1)
            if isinstance(v, list):
                return np.array(v)[None, :]
            if isinstance(v, (int, float)):
                return np.array([v])[None, :]
            return np.atleast_2d(v)

        self.obs_scale, self.coords_scale = as2d(obs_scale), as2d(coords_scale)
        self.coords = self.coords / self.coords_scale
        self.obs = (self.obs - self.obs_mean) / self.obs_scale
        self.verbose = verbose
        self.cpu_name = _processor_name()
        self.gpu_name = None
        self._device = int(device)
        try:
            import torch
            if torch.cuda.is_available():
                self.gpu_name = torch.cuda.get_device_name(self._device)
        except Exception:
            pass

        # ---- gpflow_models.py:113-157 ----
        assert kernel is not None, "kernel was not provided"
        if not isinstance(kernel, str):
            raise TypeError("B200GPRModel takes the kernel by name (a gpflow.kernels class name)")
        assert kernel in KERNEL_IDS, f"kernel '{kernel}' is not one of {sorted(KERNEL_IDS)}"
        assert mean_function is None, "B200GPRModel implements the zero prior mean path only"
        assert likelihood is None, "B200GPRModel implements the Gaussian likelihood only"
        self.kernel = kernel
        kernel_kwargs = dict(kernel_kwargs or {})
        D = self.coords.shape[1]
        self._hp = HyperParams(D, kernel_kwargs.pop("lengthscales", None), kernel_kwargs.pop("variance", 1.0),
                               1.0 if noise_variance is None else noise_variance)
        assert not kernel_kwargs, f"unsupported kernel_kwargs: {list(kernel_kwargs)}"
        self._batch = None
        self.opt_info = None

    # ---- engine plumbing ----
    def _engine(self):
        from . import get_engine
        return get_engine(self._device)

    def _get_batch(self):
        if self._batch is None:
            eng = self._engine()
            n = len(self.obs)
            self._batch = eng.make_batch(np.array([0, n], dtype=np.int64), self.coords, self.obs[:, 0],
                                         kernel=self.kernel)
        return self._batch

    # ---- parameters (base_model.py:370-439; gpflow_models.py:179-184, 339-411) ----
    @property
    def param_names(self) -> list:
        return list(PARAM_NAMES)

    def get_parameters(self, *args, return_dict=True):
        if len(args) == 0:
            args = self.param_names
        for a in args:
            assert a in self.param_names, f"cannot get parameters for: {a}, it's not in param_names: {self.param_names}"
        if return_dict:
            return {a: getattr(self, f"get_{a}")() for a in args}
        return [getattr(self, f"get_{a}")() for a in args]

    def set_parameters(self, **kwargs):
        for k, v in kwargs.items():
            assert k in self.param_names, f"cannot get parameters for: {k}, it's not in param_names: {self.param_names}"
            getattr(self, f"set_{k}")(v)

    def get_lengthscales(self):
        return self._hp.get("lengthscales")

    def get_kernel_variance(self):
        return self._hp.get("kernel_variance")

    def get_likelihood_variance(self):
        return self._hp.get("likelihood_variance")

    def set_lengthscales(self, lengthscales):
        self._hp.set("lengthscales", lengthscales)

    def set_kernel_variance(self, kernel_variance):
        self._hp.set("kernel_variance", kernel_variance)

    def set_likelihood_variance(self, likelihood_variance):
        self._hp.set("likelihood_variance", likelihood_variance)

    def set_parameter_constraints(self, constraints_dict, **kwargs):
        for k, v in constraints_dict.items():
            assert k in self.param_names, f"cannot get parameters for: {k}, it's not in param_names: {self.param_names}"
            getattr(self, f"set_{k}_constraints")(**v, **kwargs)

    def set_lengthscales_constraints(self, low, high, move_within_tol=True, tol=1e-8, scale=False,
                                     scale_magnitude=None):
        self._hp.set_constraints("lengthscales", low, high, move_within_tol=move_within_tol, tol=tol, scale=scale,
                                 scale_magnitude=scale_magnitude, coords_scale=self.coords_scale)

    def set_kernel_variance_constraints(self, low, high, move_within_tol=True, tol=1e-8, scale=False,
                                        scale_magnitude=None):
        self._hp.set_constraints("kernel_variance", low, high, move_within_tol=move_within_tol, tol=tol,
                                 scale=scale, scale_magnitude=scale_magnitude, coords_scale=self.coords_scale)

    def set_likelihood_variance_constraints(self, low, high, move_within_tol=True, tol=1e-8, scale=False,
                                            scale_magnitude=None):
        self._hp.set_constraints("likelihood_variance", low, high, move_within_tol=move_within_tol, tol=tol,
                                 scale=scale, scale_magnitude=scale_magnitude, coords_scale=self.coords_scale)

    # ---- P1 (gpflow_models.py:290-329) ----
    def optimise_parameters(self, max_iter=10_000, fixed_params=None, **opt_kwargs) -> bool:
        t0 = time.perf_counter()
        eng = self._engine()
        kind, low, high = self._hp.transforms()
        options = dict(opt_kwargs.pop("options", None) or {})
        options.update({k: opt_kwargs.pop(k) for k in ("ftol", "gtol", "maxcor", "maxls", "maxfun") if k in opt_kwargs})
        assert opt_kwargs.pop("method", "L-BFGS-B") == "L-BFGS-B", "only L-BFGS-B is implemented"
        assert not opt_kwargs, f"unsupported optimiser arguments: {list(opt_kwargs)}"
        res = eng.optimise(self._get_batch(), self._hp.theta(), kind, low, high,
                           self._hp.trainable_mask(fixed_params), maxiter=int(options.pop("maxiter", max_iter)),
                           **options)
        self._hp.set_theta(res["theta"][0].cpu().numpy())
        status = int(res["status"][0])
        from ._lib import OPT_STATUS
        self.opt_info = {"status": status, "message": OPT_STATUS[status], "nit": int(res["nit"][0]),
                         "nfev": int(res["nfev"][0]), "fun": float(res["fobj"][0])}
        success = status in (1, 2)
        if not success and self.verbose:
            print("*" * 10 + "\noptimization failed!")
        if self.verbose:
            print(f"'optimise_parameters': {time.perf_counter() - t0:.3f} seconds")
        return success

    # ---- L1 (gpflow_models.py:334-337) ----
    def get_objective_function_value(self):
        f, _ = self._engine().eval(self._get_batch(), self._hp.theta(), grad=False)
        return float(f[0])

    # ---- F1 (gpflow_models.py:186-273) ----
    def predict(self, coords, full_cov=False, apply_scale=True) -> Dict[str, np.ndarray]:
        if isinstance(coords, (pd.Series, pd.DataFrame)):
            if self.coords_col is not None:
                coords = coords[self.coords_col].values
            else:
                coords = coords.values
        if isinstance(coords, list):
            coords = np.array(coords)
        if len(coords.shape) == 1:
            coords = coords[None, :]
        assert isinstance(coords, np.ndarray), f"coords should be an ndarray (one can be converted from)"
        coords = coords.astype(self.coords.dtype)
        if apply_scale:
            coords = coords / self.coords_scale
        eng = self._engine()
        P = len(coords)
        if full_cov:
            fm, fcov = eng.predict_full_cov(self._get_batch(), self._hp.theta(), coords)
            f_cov = fcov.cpu().numpy()
            f_var = np.diag(f_cov)
            y_var = f_var + self._hp.likelihood_variance
            y_cov = f_cov.copy()
            y_cov[np.arange(P), np.arange(P)] += y_var - f_var
            out = {"f*": fm.cpu().numpy(), "f*_var": f_var, "y_var": y_var, "f*_cov": f_cov, "y_cov": y_cov}
        else:
            fm, fv, yv, _ = eng.predict(self._get_batch(), self._hp.theta(), np.array([0, P], dtype=np.int64),
                                        np.ascontiguousarray(coords))
            out = {"f*": fm.cpu().numpy(), "f*_var": fv.cpu().numpy(), "y_var": yv.cpu().numpy()}
        f_bar = self.obs_mean[:, 0]
        if len(f_bar) != len(out["f*"]):
            assert len(f_bar) == 1, \
                f"'f_bar' did not match the length of 'f*' and f_bar len is not, got: {len(f_bar)}"
            out["f_bar"] = np.repeat(f_bar, len(out["f*"]))
        else:
            out["f_bar"] = f_bar
        return out


def get_model(name):
    """Name lookup in the style of GPSat.models.get_model (models/__init__.py:3-26)."""
    if name in ("B200GPRModel", "GPflowGPRModel"):
        return B200GPRModel
    if name in ("B200SGPRModel", "GPflowSGPRModel"):
        return B200SGPRModel
    raise NotImplementedError(f"model: {name} is not implemented by gpsat_b200")


def register(gpsat_module=None):
    """Make ``"oi_model": "B200GPRModel"`` resolve inside an importable GPSat.

    GPSat.models.get_model is a hard-coded if/elif chain imported by value into three namespaces
    (models/__init__.py:3, local_experts.py:32, postprocessing.py:18); all three are wrapped.
    Without touching the reference one can equally use the dict form of ``oi_model``
    ({"path_to_model": "gpsat_b200.model", "model_name": "B200GPRModel"}, local_experts.py:319-325).
    """
    import importlib
    import sys
    mods = []
    for nm in ("GPSat.models", "GPSat.local_experts", "GPSat.postprocessing"):
        try:
            mods.append(sys.modules.get(nm) or importlib.import_module(nm))
        except Exception:      # the module may be unimportable (tensorflow / tables absent)
            continue
    for m in mods:
        orig = getattr(m, "get_model", None)
        if orig is None or getattr(orig, "_gpsat_b200", False):
            continue

        def wrapped(name, _orig=orig):
            if name == "B200GPRModel":
                return B200GPRModel
            if name == "B200SGPRModel":
                return B200SGPRModel
            return _orig(name)

        wrapped._gpsat_b200 = True
        m.get_model = wrapped
    return [m.__name__ for m in mods]


class B200SGPRModel(B200GPRModel):
    """Sparse GPR with M inducing points: drop-in for ``GPflowSGPRModel`` (gpflow_models.py:666-901).

    Inducing points follow the reference exactly: all data if N < M, else the first M rows of a
    ``np.random.shuffle`` of a copy of the (scaled) coordinates -- reproducible only if the caller
    seeds numpy's global RNG, like the reference (gpflow_models.py:809-819).
    ``get_objective_function_value`` returns +ELBO (gpflow_models.py:860-862).
    """

    def __init__(self, data=None, coords_col=None, obs_col=None, coords=None, obs=None, coords_scale=None,
                 obs_scale=None, obs_mean=None, verbose=True, *, kernel="Matern32", num_inducing_points=500,
                 kernel_kwargs=None, mean_function=None, mean_func_kwargs=None, noise_variance=None,
                 likelihood=None, **kwargs):
        super().__init__(data=data, coords_col=coords_col, obs_col=obs_col, coords=coords, obs=obs,
                         coords_scale=coords_scale, obs_scale=obs_scale, obs_mean=obs_mean, verbose=verbose,
                         kernel=kernel, kernel_kwargs=kernel_kwargs, mean_function=mean_function,
                         mean_func_kwargs=mean_func_kwargs, noise_variance=noise_variance, likelihood=likelihood,
                         **kwargs)
        assert num_inducing_points is not None, "num_inducing_points is None, must be specified for SGPR"
        if len(self.coords) < num_inducing_points:
            if verbose:
                print("number of inducing points is more than number of data points, "
                      "setting inducing points to data points...")
            self.inducing_points = self.coords
        else:
            X = self.coords.copy()
            np.random.shuffle(X)
            self.inducing_points = X[:num_inducing_points]
        self._sbatch = None

    @property
    def param_names(self) -> list:
        return list(PARAM_NAMES) + ["inducing_points"]

    def get_inducing_points(self) -> np.ndarray:
        return np.array(self.inducing_points)

    def set_inducing_points(self, inducing_points):
        self.inducing_points = np.array(inducing_points, dtype=np.float64)
        self._sbatch = None

    def _get_sbatch(self):
        if self._sbatch is None:
            eng = self._engine()
            Z = np.ascontiguousarray(self.inducing_points, dtype=np.float64)
            self._sbatch = eng.make_sgpr_batch(self._get_batch(), np.array([0, len(Z)], dtype=np.int64), Z)
        return self._sbatch

    def get_objective_function_value(self):
        f, _ = self._engine().sgpr_eval(self._get_sbatch(), self._hp.theta(), grad=False)
        return -float(f[0])

    def optimise_parameters(self, train_inducing_points=False, max_iter=10_000, fixed_params=None, **opt_kwargs):
        assert not train_inducing_points, "B200SGPRModel keeps the inducing points fixed (the reference's default)"
        t0 = time.perf_counter()
        kind, low, high = self._hp.transforms()
        options = dict(opt_kwargs.pop("options", None) or {})
        options.update({k: opt_kwargs.pop(k) for k in ("ftol", "gtol", "maxcor", "maxls", "maxfun") if k in opt_kwargs})
        assert opt_kwargs.pop("method", "L-BFGS-B") == "L-BFGS-B", "only L-BFGS-B is implemented"
        assert not opt_kwargs, f"unsupported optimiser arguments: {list(opt_kwargs)}"
        res = self._engine().sgpr_optimise(self._get_sbatch(), self._hp.theta(), kind, low, high,
                                           self._hp.trainable_mask(fixed_params),
                                           maxiter=int(options.pop("maxiter", max_iter)), **options)
        self._hp.set_theta(res["theta"][0].cpu().numpy())
        status = int(res["status"][0])
        from ._lib import OPT_STATUS
        self.opt_info = {"status": status, "message": OPT_STATUS[status], "nit": int(res["nit"][0]),
                         "nfev": int(res["nfev"][0]), "fun": float(res["fobj"][0])}
        if self.verbose:
            print(f"'optimise_parameters': {time.perf_counter() - t0:.3f} seconds")
        return status in (1, 2)

    def predict(self, coords, full_cov=False, apply_scale=True) -> Dict[str, np.ndarray]:
        assert not full_cov, "full_cov is not implemented for the sparse model"
        if isinstance(coords, (pd.Series, pd.DataFrame)):
            coords = coords[self.coords_col].values if self.coords_col is not None else coords.values
        if isinstance(coords, list):
            coords = np.array(coords)
        if len(coords.shape) == 1:
            coords = coords[None, :]
        coords = coords.astype(self.coords.dtype)
        if apply_scale:
            coords = coords / self.coords_scale
        P = len(coords)
        fm, fv, yv, _ = self._engine().sgpr_predict(self._get_sbatch(), self._hp.theta(),
                                                    np.array([0, P], dtype=np.int64), np.ascontiguousarray(coords))
        out = {"f*": fm.cpu().numpy(), "f*_var": fv.cpu().numpy(), "y_var": yv.cpu().numpy()}
        f_bar = self.obs_mean[:, 0]
        out["f_bar"] = np.repeat(f_bar, P) if len(f_bar) != P else f_bar
        return out
