"""Hyper-parameter bookkeeping on the host: values, transforms and constraints.

Mirrors the observable behaviour of the reference's GPflow wrapper (citations relative to the
reference tree):
  * defaults                      GPSat/models/gpflow_models.py:116-157  (lengthscales = 1, kernel
                                  variance 1, likelihood variance 1; gpflow positive() = softplus;
                                  Gaussian likelihood variance = softplus + 1e-6)
  * getters / setters             gpflow_models.py:339-411
  * constraints + move_within_tol gpflow_models.py:416-494, 592-628 (tfp.bijectors.Sigmoid(low, high))
The device only ever sees the flat arrays produced by ``theta()`` / ``transforms()``.
"""
from __future__ import annotations

import warnings

import numpy as np

PARAM_NAMES = ["lengthscales", "kernel_variance", "likelihood_variance"]
LIK_VAR_LOWER = 1e-6          # gpflow.likelihoods.Gaussian DEFAULT_VARIANCE_LOWER_BOUND
SOFTPLUS, SIGMOID = 0, 1      # gpsat_transforms.kind


class HyperParams:
    """lengthscales[D], kernel_variance, likelihood_variance with their bijectors."""

    def __init__(self, D, lengthscales=None, kernel_variance=1.0, likelihood_variance=1.0):
        self.D = int(D)
        ls = np.ones(self.D) if lengthscales is None else np.asarray(lengthscales, dtype=np.float64)
        self.lengthscales = np.broadcast_to(ls, (self.D,)).astype(np.float64).copy()
        self.kernel_variance = float(kernel_variance)
        self.likelihood_variance = float(likelihood_variance)
        # per parameter: (kind, low[], high[])
        self.tr = {"lengthscales": (SOFTPLUS, np.zeros(self.D), np.zeros(self.D)),
                   "kernel_variance": (SOFTPLUS, np.zeros(1), np.zeros(1)),
                   "likelihood_variance": (SOFTPLUS, np.full(1, LIK_VAR_LOWER), np.zeros(1))}

    def copy(self):
        out = HyperParams(self.D, self.lengthscales, self.kernel_variance, self.likelihood_variance)
        out.tr = {k: (v[0], v[1].copy(), v[2].copy()) for k, v in self.tr.items()}
        return out

    # ---- getters / setters (gpflow_models.py:339-411) ----
    def get(self, name):
        assert name in PARAM_NAMES, f"name: {name} not in param_names: {PARAM_NAMES}"
        if name == "lengthscales":
            return self.lengthscales.copy()
        return float(getattr(self, name))

    def set(self, name, value):
        assert name in PARAM_NAMES, f"name: {name} not in param_names: {PARAM_NAMES}"
        if name == "lengthscales":
            self.lengthscales = np.broadcast_to(np.asarray(value, dtype=np.float64), (self.D,)).copy()
            return
        if isinstance(value, np.ndarray):
            if value.ndim > 0:
                assert (len(value) == 1) & (value.ndim == 1), "expected a scalar or an array of length 1"
                value = value[0]
        elif isinstance(value, (list, tuple)):
            assert len(value) == 1
            value = value[0]
        value = float(value)
        if name == "likelihood_variance" and value < LIK_VAR_LOWER:
            warnings.warn(f"likelihood_variance {value} below the lower bound {LIK_VAR_LOWER}; clipped")
            value = LIK_VAR_LOWER
        setattr(self, name, value)

    # ---- constraints (gpflow_models.py:416-494) ----
    def set_constraints(self, name, low, high, move_within_tol=True, tol=1e-8, scale=False,
                        scale_magnitude=None, coords_scale=None):
        assert name in PARAM_NAMES, f"name: {name} not in param_names: {PARAM_NAMES}"

        def arr(v):
            if isinstance(v, (list, tuple)):
                return np.array(v, dtype=np.float64)
            if isinstance(v, (int, np.integer, float, np.floating)):
                return np.array([v], dtype=np.float64)
            return np.asarray(v, dtype=np.float64)

        low, high = arr(low), arr(high)
        assert low.ndim == 1 and high.ndim == 1, "low / high must be scalars or 1-d"
        vals = np.atleast_1d(np.array(self.get(name), dtype=np.float64))
        orig = vals.copy()
        assert len(vals) == len(low), "len of low constraint does not match param length"
        assert len(vals) == len(high), "len of high constraint does not match param length"
        assert np.all(low <= high), "all values in high constraint must be greater than low"
        if scale:
            if scale_magnitude is None:
                cs = np.atleast_2d(1.0 if coords_scale is None else coords_scale)[0, :]
                low, high = low / cs, high / cs
            else:
                low, high = low / scale_magnitude, high / scale_magnitude
        if move_within_tol:
            half_min_width = np.min(high - low) / 2
            if tol > half_min_width:
                tol = half_min_width
            m = vals > (high - tol)
            vals[m] = high[m] - tol
            m = vals < (low + tol)
            vals[m] = low[m] + tol
        if (orig != vals).any():
            self.set(name, vals if name == "lengthscales" else vals[0])
        self.tr[name] = (SIGMOID, low.copy(), high.copy())

    # ---- flat views for the device ----
    def theta(self):
        return np.concatenate([self.lengthscales, [self.kernel_variance], [self.likelihood_variance]])

    def set_theta(self, th):
        th = np.asarray(th, dtype=np.float64)
        self.lengthscales = th[:self.D].copy()
        self.kernel_variance = float(th[self.D])
        self.likelihood_variance = float(th[self.D + 1])

    def transforms(self):
        """(kind[D+2], low[D+2], high[D+2]) in theta order."""
        kind, low, high = [], [], []
        for nm in PARAM_NAMES:
            k, lo, hi = self.tr[nm]
            n = self.D if nm == "lengthscales" else 1
            kind += [k] * n
            low += list(np.broadcast_to(lo, (n,)))
            high += list(np.broadcast_to(hi, (n,)))
        return np.array(kind, dtype=np.int32), np.array(low, dtype=np.float64), np.array(high, dtype=np.float64)

    def trainable_mask(self, fixed_params=None):
        """gpflow_models.py:275-288: names in fixed_params are removed from the variable vector."""
        fixed_params = fixed_params or []
        for p in fixed_params:
            assert p in PARAM_NAMES, f"fixed parameter {p} not in param_names: {PARAM_NAMES}"
        m = np.ones(self.D + 2, dtype=np.int32)
        if "lengthscales" in fixed_params:
            m[:self.D] = 0
        if "kernel_variance" in fixed_params:
            m[self.D] = 0
        if "likelihood_variance" in fixed_params:
            m[self.D + 1] = 0
        return m
