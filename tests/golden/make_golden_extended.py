"""Extended-precision (x87 80-bit long double, 64-bit mantissa) reference for the full-size fixed-parameter checks.

    python tests/golden/make_golden_extended.py          (authoring container, ~10 minutes, numpy only)

At N ~ 2-4 k the float64 oracle (numpy/LAPACK) and the float64 CUDA path both carry rounding error of order
cond(K_y) * 2^-53 in the predictive mean and -- through the cancellation f*_var = k** - sum A^2 -- in the variance, so
a disagreement of 1e-7 between the two says nothing about which one is right.  This script evaluates rows K1 / L1 / F1
of SURVEY.md section 8a (the GPflow arithmetic reached from GPSat/models/gpflow_models.py:229-230,337) in long double
on the SAME seeded inputs as tests/test_gpu_parity.py::test_large_experts_objective_gradient_predict and stores the
results rounded to float64 (tests/golden/extended.npz).  The GPU test then requires

    |gpu - extended| <= max(1e-8 |extended|, |numpy_float64 - extended|)        per output vector (max norm)

i.e. the CUDA path meets BASELINE.json's 1e-8 or is at least as close to the true value as the float64 oracle.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
LD = np.longdouble


def inputs():
    """the generator of test_large_experts_objective_gradient_predict (seed 41)"""
    rng = np.random.default_rng(41)
    sizes = [2100, 4300]
    Xs, zs = [], []
    for n in sizes:
        xy = rng.uniform(-3e5, 3e5, (n, 2))
        t = rng.integers(18322, 18331, n).astype(np.float64)
        X = np.column_stack([xy, t])
        Xs.append(X)
        zs.append(0.1 * np.sin(X[:, 0] / 2e5) + 0.05 * np.cos(X[:, 1] / 1.5e5) + rng.normal(0, 0.05, n))
    cs = np.array([50_000.0, 50_000.0, 1.0])
    theta = np.array([[6.0, 5.0, 7.0, 0.012, 0.004], [4.0, 8.0, 5.0, 0.02, 0.003]])
    P = 150
    Xp = np.column_stack([rng.uniform(-2e5, 2e5, (P, 2)), np.full(P, 18326.0)])
    return sizes, Xs, zs, cs, theta, Xp


def matern32_ld(X, X2, ls, kvar):
    Xs, X2s = X / ls, X2 / ls
    r2 = np.zeros((len(X), len(X2)), dtype=LD)
    for d in range(X.shape[1]):
        dd = Xs[:, None, d] - X2s[None, :, d]
        r2 += dd * dd
    r = np.sqrt(np.maximum(r2, LD("1e-36")))
    s3 = np.sqrt(LD(3))
    return kvar * (1 + s3 * r) * np.exp(-s3 * r)


def chol_ld(A, nb=64):
    """in-place lower Cholesky, left-looking by block columns"""
    n = len(A)
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0:
            A[j0:, j0:j1] -= A[j0:, :j0] @ A[j0:j1, :j0].T
        for k in range(j0, j1):
            A[k, k] = np.sqrt(A[k, k] - A[k, j0:k] @ A[k, j0:k])
            if k + 1 < n:
                A[k + 1:, k] = (A[k + 1:, k] - A[k + 1:, j0:k] @ A[k, j0:k]) / A[k, k]
    return np.tril(A)


def solve_lower_ld(L, B, nb=64):
    n = len(L)
    X = np.array(B, dtype=LD, copy=True)
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0:
            X[j0:j1] -= L[j0:j1, :j0] @ X[:j0]
        for k in range(j0, j1):
            X[k] = (X[k] - L[k, j0:k] @ X[j0:k]) / L[k, k]
    return X


def main():
    assert np.finfo(LD).nmant >= 63, "needs x87 extended precision"
    from oracle import gpr
    sizes, Xs, zs, cs, theta, Xp = inputs()
    out = {}
    for e, n in enumerate(sizes):
        t0 = time.time()
        X = (Xs[e] / cs).astype(LD)
        y = zs[e].astype(LD)
        xp = (Xp / cs).astype(LD)
        ls, kvar, nvar = theta[e, :3].astype(LD), LD(theta[e, 3]), LD(theta[e, 4])
        K = matern32_ld(X, X, ls, kvar)
        K[np.diag_indices(n)] += nvar
        L = chol_ld(K)
        a = solve_lower_ld(L, y[:, None])[:, 0]
        A = solve_lower_ld(L, matern32_ld(X, xp, ls, kvar))
        f = 0.5 * (a @ a) + np.sum(np.log(np.diag(L))) + 0.5 * n * np.log(2 * LD(np.pi))   # np.pi is only float64:
        f = 0.5 * (a @ a) + np.sum(np.log(np.diag(L))) + 0.5 * n * LD("1.8378770664093454835606594728112353")
        mean = A.T @ a
        fvar = kvar - np.sum(A * A, axis=0)
        # the float64 oracle on the same inputs, for the record (the test recomputes it)
        fr = -gpr.lml(Xs[e] / cs, zs[e], theta[e, :3], theta[e, 3], theta[e, 4])
        m64, v64, _ = gpr.predict(Xs[e] / cs, zs[e], Xp / cs, theta[e, :3], theta[e, 3], theta[e, 4])
        out[f"f_{n}"] = np.float64(f)
        out[f"mean_{n}"] = mean.astype(np.float64)
        out[f"fvar_{n}"] = fvar.astype(np.float64)
        em = np.abs(m64 - out[f"mean_{n}"]).max() / np.abs(out[f"mean_{n}"]).max()
        ev = np.abs(v64 - out[f"fvar_{n}"]).max() / np.abs(out[f"fvar_{n}"]).max()
        evr = (np.abs(v64 - out[f"fvar_{n}"]) / np.abs(out[f"fvar_{n}"])).max()
        print(f"N={n}: {time.time() - t0:.0f} s; float64 oracle vs extended: objective {abs(fr - float(f)) / abs(float(f)):.2e}, "
              f"mean {em:.2e} (max norm), variance {ev:.2e} (max norm) / {evr:.2e} (elementwise)", flush=True)
        out[f"oracle64_err_{n}"] = np.array([abs(fr - float(f)) / abs(float(f)), em, ev, evr])
    np.savez(os.path.join(HERE, "extended.npz"), **out)


if __name__ == "__main__":
    main()
