"""Exact-GPR oracle (numpy/scipy float64).  TEST INFRASTRUCTURE ONLY.

Restates, per SURVEY.md section 8a:
  K1  gpflow.kernels.{Matern12,Matern32,Matern52,SquaredExponential} reached from
      GPSat/models/gpflow_models.py:116-135,153
  L1  gpflow.models.GPR.log_marginal_likelihood  (gpflow_models.py:337)
  G1  d(-LML)/d(theta) (TF autodiff in gpflow.optimizers.Scipy, gpflow_models.py:317-321),
      here in analytic form 0.5*tr((K^-1 - aa^T) dK/dtheta)
  F1  gpflow.models.GPR.predict_f / predict_y  (gpflow_models.py:229-230)
  M1  BaseGPRModel.__init__ data preparation (GPSat/models/base_model.py:134-245)
  M3  getters/setters (gpflow_models.py:334-411)
  M4  constraints + tfp Sigmoid bijector (gpflow_models.py:416-494,592-628)
  P1  scipy L-BFGS-B through gpflow.optimizers.Scipy (gpflow_models.py:290-329)
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt

SQRT3 = np.sqrt(3.0)
SQRT5 = np.sqrt(5.0)
LOG2PI = np.log(2.0 * np.pi)

KERNEL_IDS = {"Matern32": 0, "Matern52": 1, "Matern12": 2, "Exponential": 2,
              "RBF": 3, "SquaredExponential": 3}
LIK_VAR_LOWER = 1e-6  # gpflow.likelihoods.Gaussian DEFAULT_VARIANCE_LOWER_BOUND


# ----------------------------------------------------------------------------
# K1: kernels
# ----------------------------------------------------------------------------
def scaled_sqdist(X, X2, ls, form="direct"):
    """r^2 between rows of X/ls and X2/ls.

    form="direct": sum_d ((x_d - x'_d)/l_d)^2  (what the CUDA path computes)
    form="gpflow": -2 X~ X2~^T + |x~|^2 + |x2~|^2 (gpflow.utilities.ops.square_distance)
    """
    Xs = X / ls
    X2s = Xs if X2 is None else X2 / ls
    if form == "direct":
        d = Xs[:, None, :] - X2s[None, :, :]
        r2 = np.zeros(d.shape[:2])
        for k in range(d.shape[2]):  # fixed summation order d=0,1,2 (matches device)
            r2 = r2 + d[:, :, k] * d[:, :, k]
        return r2
    elif form == "gpflow":
        xs = np.sum(Xs * Xs, axis=-1)
        x2s = np.sum(X2s * X2s, axis=-1)
        return -2.0 * Xs @ X2s.T + xs[:, None] + x2s[None, :]
    raise ValueError(form)


def k_of_r2(r2, variance, kernel):
    """Stationary kernel value from r^2 (gpflow K_r / K_r2)."""
    kid = KERNEL_IDS[kernel]
    if kid == 3:
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))
    if kid == 0:
        return variance * (1.0 + SQRT3 * r) * np.exp(-SQRT3 * r)
    if kid == 1:
        return variance * (1.0 + SQRT5 * r + 5.0 / 3.0 * (r * r)) * np.exp(-SQRT5 * r)
    return variance * np.exp(-r)


def h_of_r2(r2, variance, kernel):
    """h(r) such that dk/dl_d = h(r) * delta_d^2 / l_d^3."""
    kid = KERNEL_IDS[kernel]
    if kid == 3:
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))
    if kid == 0:
        return 3.0 * variance * np.exp(-SQRT3 * r)
    if kid == 1:
        return variance * (5.0 / 3.0) * (1.0 + SQRT5 * r) * np.exp(-SQRT5 * r)
    # Matern12: gpflow clamps r^2 at 1e-36, so d/dr2 is zero below the clamp
    return np.where(r2 > 1e-36, variance * np.exp(-r) / r, 0.0)


def kh_of_r2(r2, variance, kernel):
    """(k, h) sharing the exponential (same values as k_of_r2 / h_of_r2)."""
    kid = KERNEL_IDS[kernel]
    if kid == 3:
        k = variance * np.exp(-0.5 * r2)
        return k, k.copy()
    r = np.sqrt(np.maximum(r2, 1e-36))
    if kid == 0:
        e = np.exp(-SQRT3 * r)
        return variance * (1.0 + SQRT3 * r) * e, 3.0 * variance * e
    if kid == 1:
        e = np.exp(-SQRT5 * r)
        return variance * (1.0 + SQRT5 * r + 5.0 / 3.0 * (r * r)) * e, variance * (5.0 / 3.0) * (1.0 + SQRT5 * r) * e
    k = variance * np.exp(-r)
    return k, np.where(r2 > 1e-36, k / r, 0.0)


def kernel_matrix(X, X2, ls, variance, kernel="Matern32", form="direct"):
    return k_of_r2(scaled_sqdist(X, X2, ls, form), variance, kernel)


# ----------------------------------------------------------------------------
# L1 / G1: log marginal likelihood and gradient wrt constrained theta
# ----------------------------------------------------------------------------
def lml(X, y, ls, kvar, nvar, kernel="Matern32", form="direct"):
    """GPR.log_marginal_likelihood: -0.5 a'a - sum log L_ii - N/2 log 2pi."""
    n = X.shape[0]
    K = kernel_matrix(X, None, ls, kvar, kernel, form)
    K[np.diag_indices(n)] += nvar
    L = np.linalg.cholesky(K)
    a = sla.solve_triangular(L, y, lower=True)
    return -0.5 * float(a @ a) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * LOG2PI


def neg_lml_and_grad(X, y, ls, kvar, nvar, kernel="Matern32"):
    """(-LML, d(-LML)/d[ls..., kvar, nvar]) using the direct distance form.

    LAPACK potrf / potri (threaded) so that the CPU baseline timed by bench.py is not handicapped by
    an explicit triangular inverse; raises LinAlgError when K is not positive definite.
    """
    n, D = X.shape
    Xs = X / ls
    d2 = []
    r2 = np.zeros((n, n))
    for d in range(D):                      # fixed summation order d = 0, 1, 2 (matches scaled_sqdist / the device)
        dd = Xs[:, None, d] - Xs[None, :, d]
        dd *= dd
        d2.append(dd)
        r2 += dd
    Kf, h = kh_of_r2(r2, kvar, kernel)
    K = Kf.copy()
    K[np.diag_indices(n)] += nvar
    L, info = sla.lapack.dpotrf(K, lower=1, clean=1, overwrite_a=1)
    if info != 0:
        raise np.linalg.LinAlgError("matrix is not positive definite")
    a = sla.solve_triangular(L, y, lower=True)
    alpha = sla.solve_triangular(L, a, lower=True, trans="T")
    Kinv, info = sla.lapack.dpotri(L, lower=1, overwrite_c=0)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri failed")
    Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
    f = 0.5 * float(a @ a) + float(np.sum(np.log(np.diag(L)))) + 0.5 * n * LOG2PI
    W = Kinv
    W -= np.outer(alpha, alpha)
    g = np.zeros(D + 2)
    h *= W
    for d in range(D):
        g[d] = 0.5 * float(np.einsum("ij,ij->", h, d2[d])) / ls[d]
    g[D] = 0.5 * float(np.einsum("ij,ij->", W, Kf)) / kvar
    g[D + 1] = 0.5 * np.trace(W)
    return f, g


# ----------------------------------------------------------------------------
# F1: prediction
# ----------------------------------------------------------------------------
def predict(X, y, Xs, ls, kvar, nvar, kernel="Matern32", form="direct", full_cov=False):
    """gpflow predict_f (+ predict_y variance): returns f*, f*_var, y_var [, f*_cov]."""
    n = X.shape[0]
    K = kernel_matrix(X, None, ls, kvar, kernel, form)
    K[np.diag_indices(n)] += nvar
    L = np.linalg.cholesky(K)
    Kxs = kernel_matrix(X, Xs, ls, kvar, kernel, form)
    A = sla.solve_triangular(L, Kxs, lower=True)
    a = sla.solve_triangular(L, y, lower=True)
    mean = A.T @ a
    if full_cov:
        Kss = kernel_matrix(Xs, None, ls, kvar, kernel, form)
        fcov = Kss - A.T @ A
        fvar = np.diag(fcov).copy()
        return mean, fvar, fvar + nvar, fcov
    fvar = kvar - np.sum(A * A, axis=0)
    return mean, fvar, fvar + nvar


# ----------------------------------------------------------------------------
# bijectors (tfp.bijectors.Sigmoid(low, high), Softplus, Shift∘Softplus)
# ----------------------------------------------------------------------------
def sigmoid(u):
    u = np.asarray(u, dtype=np.float64)
    out = np.empty_like(u)
    pos = u >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-u[pos]))
    e = np.exp(u[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def softplus(u):
    u = np.asarray(u, dtype=np.float64)
    return np.maximum(u, 0.0) + np.log1p(np.exp(-np.abs(u)))


def softplus_inv(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


class Transform:
    """theta = fwd(u).  kind 0: softplus + shift;  kind 1: low + (high-low)*sigmoid(u)."""

    def __init__(self, kind, low=0.0, high=0.0):
        self.kind = kind
        self.low = np.atleast_1d(np.asarray(low, dtype=np.float64))
        self.high = np.atleast_1d(np.asarray(high, dtype=np.float64))

    def fwd(self, u):
        if self.kind == 0:
            return softplus(u) + self.low
        return self.low + (self.high - self.low) * sigmoid(u)

    def inv(self, th):
        th = np.atleast_1d(np.asarray(th, dtype=np.float64))
        if self.kind == 0:
            return softplus_inv(th - self.low)
        x = (th - self.low) / (self.high - self.low)
        return np.log(x) - np.log1p(-x)

    def dfwd(self, u):
        s = sigmoid(u)
        if self.kind == 0:
            return s
        return (self.high - self.low) * s * (1.0 - s)


# ----------------------------------------------------------------------------
# The model: mirrors GPflowGPRModel's observable behaviour
# ----------------------------------------------------------------------------
class OracleGPRModel:
    """CPU restatement of GPSat.models.gpflow_models.GPflowGPRModel (no mean function)."""

    def __init__(self, data=None, coords_col=None, obs_col=None, coords=None, obs=None,
                 coords_scale=None, obs_scale=None, obs_mean=None, verbose=False, *,
                 kernel="Matern32", kernel_kwargs=None, mean_function=None,
                 noise_variance=None, r2_form="direct", **kwargs):
        # --- base_model.py:134-245 ---
        if data is not None:
            assert coords_col is not None, "data was provided, but coord_col was not"
            assert obs_col is not None, "data was provided, but obs_col was not"
            if isinstance(coords_col, str):
                coords_col = [coords_col]
            if isinstance(obs_col, str):
                obs_col = [obs_col]
            self.obs = np.array(data.loc[:, obs_col].values, dtype=np.float64)
            self.coords = np.array(data.loc[:, coords_col].values, dtype=np.float64)
        else:
            assert obs is not None and coords is not None
            obs = np.array(obs, dtype=np.float64)
            coords = np.array(coords, dtype=np.float64)
            if obs.ndim == 1:
                obs = obs[:, None]
            if coords.ndim == 1:
                coords = coords[:, None]
            assert len(obs) == len(coords), "obs and coords lengths don't match "
            self.obs, self.coords = obs, coords
            if coords_col is None:
                coords_col = list(range(coords.shape[1]))
            if obs_col is None:
                obs_col = [0]
        self.coords_col, self.obs_col = coords_col, obs_col
        assert not np.isnan(self.coords).any(), "nans found in coords"
        assert not np.isnan(self.obs).any(), "nans found in obs"
        if isinstance(obs_mean, str) and obs_mean == "local":
            obs_mean = np.mean(self.obs, axis=0)
        else:  # base_model.py:199-200: anything else becomes 0
            obs_mean = np.array([0])[None, :]
        self.obs_mean = np.atleast_2d(obs_mean)

        def _as2d(v):
            if v is None:
                return np.atleast_2d(1)
            if isinstance(v, list):
                return np.array(v)[None, :]
            if isinstance(v, (int, float)):
                return np.array([v])[None, :]
            return np.atleast_2d(v)

        self.obs_scale = _as2d(obs_scale)
        self.coords_scale = _as2d(coords_scale)
        self.coords = self.coords / self.coords_scale
        self.obs = (self.obs - self.obs_mean) / self.obs_scale
        self.gpu_name, self.cpu_name = None, "oracle-cpu"

        # --- gpflow_models.py:113-157 ---
        assert kernel is not None, "kernel was not provided"
        assert kernel in KERNEL_IDS, f"kernel {kernel} not supported by oracle"
        assert mean_function is None, "oracle restates the zero-mean path only"
        kernel_kwargs = dict(kernel_kwargs or {})
        D = self.coords.shape[1]
        ls = kernel_kwargs.get("lengthscales", np.ones(D))
        self.kernel = kernel
        self.r2_form = r2_form
        self.ls = np.broadcast_to(np.asarray(ls, dtype=np.float64), (D,)).copy()
        self.kvar = float(kernel_kwargs.get("variance", 1.0))
        self.nvar = 1.0 if noise_variance is None else float(noise_variance)
        # default gpflow transforms: positive() = softplus; likelihood = softplus + 1e-6
        self.tr = {"lengthscales": Transform(0, np.zeros(D)),
                   "kernel_variance": Transform(0, 0.0),
                   "likelihood_variance": Transform(0, LIK_VAR_LOWER)}
        self.opt_result = None

    # ---- params ----
    HYPERS = ["lengthscales", "kernel_variance", "likelihood_variance"]

    @property
    def param_names(self):
        return list(self.HYPERS)

    def get_lengthscales(self):
        return self.ls.copy()

    def get_kernel_variance(self):
        return float(self.kvar)

    def get_likelihood_variance(self):
        return float(self.nvar)

    def set_lengthscales(self, v):
        self.ls = np.broadcast_to(np.asarray(v, dtype=np.float64), self.ls.shape).copy()

    def set_kernel_variance(self, v):
        if isinstance(v, np.ndarray):
            assert (len(v) == 1) & (v.ndim == 1)
            v = v[0]
        self.kvar = float(v)

    def set_likelihood_variance(self, v):
        if isinstance(v, np.ndarray):
            assert (len(v) == 1) & (v.ndim == 1)
            v = v[0]
        if v < LIK_VAR_LOWER:  # gpflow_models.py:404-409
            v = LIK_VAR_LOWER
        self.nvar = float(v)

    def get_parameters(self, *args, return_dict=True):
        if len(args) == 0:
            args = self.param_names
        for a in args:
            assert a in self.param_names
        if return_dict:
            return {a: getattr(self, f"get_{a}")() for a in args}
        return [getattr(self, f"get_{a}")() for a in args]

    def set_parameters(self, **kwargs):
        for k, v in kwargs.items():
            assert k in self.param_names
            getattr(self, f"set_{k}")(v)

    # ---- constraints (gpflow_models.py:416-494) ----
    def _set_param_constraints(self, name, low, high, move_within_tol=True, tol=1e-8,
                               scale=False, scale_magnitude=None):
        if isinstance(low, (list, tuple)):
            low = np.array(low, dtype=np.float64)
        elif isinstance(low, (int, np.integer, float)):
            low = np.array([low], dtype=np.float64)
        if isinstance(high, (list, tuple)):
            high = np.array(high, dtype=np.float64)
        elif isinstance(high, (int, np.integer, float)):
            high = np.array([high], dtype=np.float64)
        assert low.ndim == 1 and high.ndim == 1
        vals = np.atleast_1d(np.array(getattr(self, f"get_{name}")(), dtype=np.float64))
        orig = vals.copy()
        assert len(vals) == len(low), "len of low constraint does not match param length"
        assert len(vals) == len(high), "len of high constraint does not match param length"
        assert np.all(low <= high), "all values in high constraint must be greater than low"
        if scale:
            if scale_magnitude is None:
                low = low / self.coords_scale[0, :]
                high = high / self.coords_scale[0, :]
            else:
                low = low / scale_magnitude
                high = high / scale_magnitude
        if move_within_tol:
            half_min_width = np.min(high - low) / 2
            if tol > half_min_width:
                tol = half_min_width
            m = vals > (high - tol)
            vals[m] = high[m] - tol
            m = vals < (low + tol)
            vals[m] = low[m] + tol
        if (orig != vals).any():
            if name == "lengthscales":
                self.ls = vals
            elif name == "kernel_variance":
                self.kvar = float(vals[0])
            else:
                self.nvar = float(vals[0])
        self.tr[name] = Transform(1, low, high)

    def set_lengthscales_constraints(self, low, high, **kw):
        self._set_param_constraints("lengthscales", low, high, **kw)

    def set_kernel_variance_constraints(self, low, high, **kw):
        self._set_param_constraints("kernel_variance", low, high, **kw)

    def set_likelihood_variance_constraints(self, low, high, **kw):
        self._set_param_constraints("likelihood_variance", low, high, **kw)

    def set_parameter_constraints(self, constraints_dict, **kwargs):
        for k, v in constraints_dict.items():
            assert k in self.param_names
            getattr(self, f"set_{k}_constraints")(**v, **kwargs)

    # ---- objective ----
    def get_objective_function_value(self):
        return -lml(self.coords, self.obs[:, 0], self.ls, self.kvar, self.nvar, self.kernel,
                    self.r2_form)

    def _theta(self):
        return np.concatenate([self.ls, [self.kvar], [self.nvar]])

    def _transforms_flat(self):
        """per-element (kind, low, high) arrays over [ls..., kvar, nvar]."""
        kind, low, high = [], [], []
        for nm in self.HYPERS:
            t = self.tr[nm]
            n = len(self.ls) if nm == "lengthscales" else 1
            kind += [t.kind] * n
            low += list(np.broadcast_to(t.low, (n,)))
            high += list(np.broadcast_to(t.high if t.kind == 1 else np.zeros(n), (n,)))
        return np.array(kind), np.array(low, dtype=np.float64), np.array(high, dtype=np.float64)

    def unconstrained(self):
        return np.concatenate([self.tr[nm].inv(np.atleast_1d(getattr(self, f"get_{nm}")()))
                               for nm in self.HYPERS])

    def objective_u(self, u_free, free_mask, u_all):
        """f(u), df/du on the trainable subset (what gpflow.optimizers.Scipy hands scipy)."""
        u = u_all.copy()
        u[free_mask] = u_free
        D = len(self.ls)
        th, dth = [], []
        off = 0
        for nm in self.HYPERS:
            n = D if nm == "lengthscales" else 1
            th.append(self.tr[nm].fwd(u[off:off + n]))
            dth.append(self.tr[nm].dfwd(u[off:off + n]))
            off += n
        th = np.concatenate(th)
        dth = np.concatenate(dth)
        try:
            f, g = neg_lml_and_grad(self.coords, self.obs[:, 0], th[:D], th[D], th[D + 1],
                                    self.kernel)
        except np.linalg.LinAlgError:
            return np.inf, np.full(free_mask.sum(), np.nan)
        return f, (g * dth)[free_mask]

    def optimise_parameters(self, max_iter=10_000, fixed_params=None, optimiser="scipy", **opt_kwargs):
        """gpflow_models.py:290-329: scipy L-BFGS-B on the unconstrained trainables."""
        fixed_params = fixed_params or []
        D = len(self.ls)
        free = np.ones(D + 2, dtype=bool)
        if "lengthscales" in fixed_params:
            free[:D] = False
        if "kernel_variance" in fixed_params:
            free[D] = False
        if "likelihood_variance" in fixed_params:
            free[D + 1] = False
        u_all = self.unconstrained()
        fun = lambda uf: self.objective_u(uf, free, u_all)
        if optimiser == "scipy":
            res = sopt.minimize(fun, u_all[free], jac=True, method="L-BFGS-B",
                                options=dict(maxiter=max_iter), **opt_kwargs)
            success, xf = bool(res.success), res.x
        else:
            from .lbfgs import minimize_lbfgs
            res = minimize_lbfgs(fun, u_all[free], maxiter=max_iter)
            success, xf = res["success"], res["x"]
        self.opt_result = res
        u_all[free] = xf
        off = 0
        for nm in self.HYPERS:
            n = D if nm == "lengthscales" else 1
            v = self.tr[nm].fwd(u_all[off:off + n])
            if nm == "lengthscales":
                self.ls = v
            elif nm == "kernel_variance":
                self.kvar = float(v[0])
            else:
                self.nvar = float(v[0])
            off += n
        return success

    # ---- predict (gpflow_models.py:186-273) ----
    def predict(self, coords, full_cov=False, apply_scale=True):
        import pandas as pd
        if isinstance(coords, (pd.Series, pd.DataFrame)):
            coords = coords[self.coords_col].values if self.coords_col is not None else coords.values
        if isinstance(coords, list):
            coords = np.array(coords)
        if coords.ndim == 1:
            coords = coords[None, :]
        coords = coords.astype(self.coords.dtype)
        if apply_scale:
            coords = coords / self.coords_scale
        res = predict(self.coords, self.obs[:, 0], coords, self.ls, self.kvar, self.nvar,
                      self.kernel, self.r2_form, full_cov=full_cov)
        out = {"f*": res[0], "f*_var": res[1], "y_var": res[2]}
        if full_cov:
            out["f*_cov"] = res[3]
            y_cov = res[3].copy()
            y_cov[np.diag_indices(len(y_cov))] += (res[2] - res[1])
            out["y_cov"] = y_cov
        f_bar = self.obs_mean[:, 0]
        out["f_bar"] = np.repeat(f_bar, len(out["f*"])) if len(f_bar) != len(out["f*"]) else f_bar
        return out
