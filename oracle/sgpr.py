"""Sparse-GPR oracle (numpy/scipy float64).  TEST INFRASTRUCTURE ONLY.

Restates row SG1 of SURVEY.md section 8a: GPSat/models/gpflow_models.py:666-901 (GPflowSGPRModel) and the
gpflow-2.9 arithmetic it calls (gpflow.models.SGPR: Titsias' collapsed bound ``elbo`` and
``predict_f``; Kuu carries gpflow's default jitter 1e-6).  gpflow itself cannot be imported here, so
the formulas are restated; they are pinned by the reference's own test
tests/test_localexperts.py:229-251 (M = N = 50 must reproduce sklearn's exact GPR to 1e-4, KAT-2)
and the gradient by finite differences.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt

from .gpr import (KERNEL_IDS, LOG2PI, OracleGPRModel, h_of_r2, k_of_r2, scaled_sqdist)

JITTER = 1e-6  # gpflow.config.default_jitter()


def _parts(X, y, Z, ls, kvar, nvar, kernel):
    N, M = X.shape[0], Z.shape[0]
    Kuu = k_of_r2(scaled_sqdist(Z, None, ls), kvar, kernel) + JITTER * np.eye(M)
    Kuf = k_of_r2(scaled_sqdist(Z, X, ls), kvar, kernel)
    L = np.linalg.cholesky(Kuu)
    sigma = np.sqrt(nvar)
    A = sla.solve_triangular(L, Kuf, lower=True) / sigma
    B = A @ A.T + np.eye(M)
    LB = np.linalg.cholesky(B)
    c = sla.solve_triangular(LB, A @ y, lower=True) / sigma
    return N, M, Kuu, Kuf, L, A, B, LB, c


def elbo(X, y, Z, ls, kvar, nvar, kernel="Matern32"):
    """gpflow.models.SGPR.elbo (zero mean function)."""
    N, M, Kuu, Kuf, L, A, B, LB, c = _parts(X, y, Z, ls, kvar, nvar, kernel)
    bound = -0.5 * N * LOG2PI
    bound -= np.sum(np.log(np.diag(LB)))
    bound -= 0.5 * N * np.log(nvar)
    bound -= 0.5 * float(y @ y) / nvar
    bound += 0.5 * float(c @ c)
    bound -= 0.5 * N * kvar / nvar          # Kdiag of a stationary kernel
    bound += 0.5 * np.sum(A * A)            # tr(A A')
    return float(bound)


def neg_elbo_and_grad(X, y, Z, ls, kvar, nvar, kernel="Matern32"):
    """(-ELBO, d(-ELBO)/d[ls..., kvar, nvar]) -- what gpflow's training_loss + autodiff hand scipy.

    With beta = 1/nvar, Sigma = Kuu + beta Kuf Kfu, w = Sigma^-1 Kuf y:
      dF/dKuf = [beta (Kuu^-1 - Sigma^-1) - beta^3 w w'] Kuf + beta^2 w y'
      dF/dKuu = 1/2 (Kuu^-1 - Sigma^-1) - 1/2 beta^2 w w' - 1/2 beta Kuu^-1 Kuf Kfu Kuu^-1
    contracted with the elementwise kernel derivatives.
    """
    N, D = X.shape
    M = Z.shape[0]
    beta = 1.0 / nvar
    r2uu = scaled_sqdist(Z, None, ls)
    r2uf = scaled_sqdist(Z, X, ls)
    Kuu0 = k_of_r2(r2uu, kvar, kernel)
    Kuf = k_of_r2(r2uf, kvar, kernel)
    Kuu = Kuu0 + JITTER * np.eye(M)
    Sigma = Kuu + beta * Kuf @ Kuf.T
    Kuu_inv = np.linalg.inv(Kuu)
    Sig_inv = np.linalg.inv(Sigma)
    v = Kuf @ y
    w = Sig_inv @ v
    F = elbo(X, y, Z, ls, kvar, nvar, kernel)
    C = beta * (Kuu_inv - Sig_inv) - beta ** 3 * np.outer(w, w)
    Guf = C @ Kuf + beta ** 2 * np.outer(w, y)
    P = Kuf @ Kuf.T
    Guu = 0.5 * (Kuu_inv - Sig_inv) - 0.5 * beta ** 2 * np.outer(w, w) - 0.5 * beta * Kuu_inv @ P @ Kuu_inv
    g = np.zeros(D + 2)
    huu, huf = h_of_r2(r2uu, kvar, kernel), h_of_r2(r2uf, kvar, kernel)
    for d in range(D):
        duu = Z[:, None, d] - Z[None, :, d]
        duf = Z[:, None, d] - X[None, :, d]
        g[d] = (np.sum(Guu * huu * duu * duu) + np.sum(Guf * huf * duf * duf)) / ls[d] ** 3
    g[D] = (np.sum(Guu * Kuu0) + np.sum(Guf * Kuf)) / kvar - 0.5 * beta * N
    # d/d nvar = -beta^2 d/d beta
    dF_dbeta = (0.5 * N / beta - 0.5 * np.sum(Sig_inv * P) - 0.5 * float(y @ y) + beta * float(v @ w)
                - 0.5 * beta ** 2 * float(w @ P @ w) - 0.5 * N * kvar + 0.5 * np.sum(Kuu_inv * P))
    g[D + 1] = -beta ** 2 * dF_dbeta
    return -F, -g


def predict(X, y, Z, Xs, ls, kvar, nvar, kernel="Matern32"):
    """gpflow.models.SGPR.predict_f (full_cov=False) + predict_y variance."""
    N, M, Kuu, Kuf, L, A, B, LB, c = _parts(X, y, Z, ls, kvar, nvar, kernel)
    Kus = k_of_r2(scaled_sqdist(Z, Xs, ls), kvar, kernel)
    t1 = sla.solve_triangular(L, Kus, lower=True)
    t2 = sla.solve_triangular(LB, t1, lower=True)
    mean = t2.T @ c
    var = kvar + np.sum(t2 * t2, axis=0) - np.sum(t1 * t1, axis=0)
    return mean, var, var + nvar


class OracleSGPRModel(OracleGPRModel):
    """CPU restatement of GPSat.models.gpflow_models.GPflowSGPRModel."""

    def __init__(self, *args, num_inducing_points=500, inducing_points=None, **kwargs):
        super().__init__(*args, **kwargs)
        assert num_inducing_points is not None, "num_inducing_points is None, must be specified for SGPR"
        if inducing_points is not None:
            self.inducing_points = np.array(inducing_points, dtype=np.float64)
        elif len(self.coords) < num_inducing_points:
            self.inducing_points = self.coords.copy()
        else:   # gpflow_models.py:809-819: global numpy RNG shuffle of a copy, first M rows
            Xc = self.coords.copy()
            np.random.shuffle(Xc)
            self.inducing_points = Xc[:num_inducing_points]

    @property
    def param_names(self):
        return super().param_names + ["inducing_points"]

    def get_inducing_points(self):
        return self.inducing_points.copy()

    def set_inducing_points(self, v):
        self.inducing_points = np.array(v, dtype=np.float64)

    def get_objective_function_value(self):
        """+ELBO (sign differs from the exact model, gpflow_models.py:860-862)."""
        return elbo(self.coords, self.obs[:, 0], self.inducing_points, self.ls, self.kvar, self.nvar, self.kernel)

    def objective_u(self, u_free, free_mask, u_all):
        u = u_all.copy()
        u[free_mask] = u_free
        D = len(self.ls)
        th, dth, off = [], [], 0
        for nm in super().param_names:
            n = D if nm == "lengthscales" else 1
            th.append(self.tr[nm].fwd(u[off:off + n]))
            dth.append(self.tr[nm].dfwd(u[off:off + n]))
            off += n
        th, dth = np.concatenate(th), np.concatenate(dth)
        try:
            f, g = neg_elbo_and_grad(self.coords, self.obs[:, 0], self.inducing_points, th[:D], th[D], th[D + 1],
                                     self.kernel)
        except np.linalg.LinAlgError:
            return np.inf, np.full(free_mask.sum(), np.nan)
        return f, (g * dth)[free_mask]

    def unconstrained(self):
        return np.concatenate([self.tr[nm].inv(np.atleast_1d(getattr(self, f"get_{nm}")()))
                               for nm in super().param_names])

    def get_parameters(self, *args, return_dict=True):
        if len(args) == 0:
            args = self.param_names
        if return_dict:
            return {a: getattr(self, f"get_{a}")() for a in args}
        return [getattr(self, f"get_{a}")() for a in args]

    def optimise_parameters(self, train_inducing_points=False, max_iter=10_000, fixed_params=None, **opt_kwargs):
        assert not train_inducing_points, "the oracle keeps the inducing points fixed (the reference's default)"
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
        res = sopt.minimize(lambda uf: self.objective_u(uf, free, u_all), u_all[free], jac=True, method="L-BFGS-B",
                            options=dict(maxiter=max_iter), **opt_kwargs)
        self.opt_result = res
        u_all[free] = res.x
        off = 0
        for nm in super().param_names:
            n = D if nm == "lengthscales" else 1
            v = self.tr[nm].fwd(u_all[off:off + n])
            if nm == "lengthscales":
                self.ls = v
            elif nm == "kernel_variance":
                self.kvar = float(v[0])
            else:
                self.nvar = float(v[0])
            off += n
        return bool(res.success)

    def predict(self, coords, full_cov=False, apply_scale=True):
        assert not full_cov
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
        m, v, yv = predict(self.coords, self.obs[:, 0], self.inducing_points, coords, self.ls, self.kvar, self.nvar,
                           self.kernel)
        out = {"f*": m, "f*_var": v, "y_var": yv}
        f_bar = self.obs_mean[:, 0]
        out["f_bar"] = np.repeat(f_bar, len(m)) if len(f_bar) != len(m) else f_bar
        return out
