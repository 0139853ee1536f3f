"""CPU tests of the C-ABI library: loads, exports every symbol of include/gpsat_b200.h, and the
host entry of the device L-BFGS state machine reproduces scipy's L-BFGS-B trajectory."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.optimize as sopt

from gpsat_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gpsat_b200.h")).read()
    names = set(re.findall(r"\b(gpsat_[a-z_0-9]+)\s*\(", hdr))
    assert names, "no declarations found"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(_lib.exported_symbols()) <= names
    assert lib.gpsat_version() >= 100


def test_struct_layouts():
    assert C.sizeof(_lib.SelTerm) == 56
    assert C.sizeof(_lib.SelSpec) == 8 + 8 * 56
    assert C.sizeof(_lib.OptOptions) == 32


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.gpsat_create(C.byref(h), 0, 0)
    assert rc != 0
    assert b"CUDA" in lib.gpsat_last_error() or rc == -3


def _run_host_lbfgs(lib, fun, x0, **kw):
    o = _lib.OptOptions()
    lib.gpsat_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    st = (C.c_char * lib.gpsat_lbfgs_state_bytes())()
    x = np.array(x0, dtype=np.float64)
    lib.gpsat_lbfgs_init_host(st, x.ctypes.data, len(x))
    nit, nfev = C.c_int(), C.c_int()
    status = 0
    f = None
    while status == 0:
        f, g = fun(x)
        g = np.ascontiguousarray(g, dtype=np.float64)
        xn = np.empty_like(x)
        status = lib.gpsat_lbfgs_tell_host(st, C.byref(o), float(f), g.ctypes.data, xn.ctypes.data,
                                           C.byref(nit), C.byref(nfev))
        x = xn
    return x, status, nit.value, nfev.value


def test_device_lbfgs_code_matches_scipy_on_host(lib):
    def ros(x):
        return sopt.rosen(x), sopt.rosen_der(x)
    for x0 in ([-1.2, 1.0], [-1.2, 1, 0.5, 2, -1], [3.0, -2.0, 0.1, 0.7, 1.5, -0.5]):
        r = sopt.minimize(ros, np.array(x0), jac=True, method="L-BFGS-B")
        x, status, nit, nfev = _run_host_lbfgs(lib, ros, x0)
        assert status in (1, 2)
        assert (nit, nfev) == (r.nit, r.nfev)
        np.testing.assert_allclose(x, r.x, rtol=1e-9, atol=1e-12)


def test_device_lbfgs_gpr_objective_matches_oracle(lib):
    from oracle import gpr
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 6, (150, 3))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(150)
    m = gpr.OracleGPRModel(coords=X.copy(), obs=y.copy(), obs_mean="local")
    m.set_parameter_constraints({"lengthscales": {"low": [1e-8] * 3, "high": [12, 12, 9]},
                                 "likelihood_variance": {"low": 0.00125, "high": 0.01}},
                                move_within_tol=True, tol=1e-2)
    free = np.ones(5, dtype=bool)
    u0 = m.unconstrained()
    fun = lambda u: m.objective_u(u, free, u0)
    r = sopt.minimize(fun, u0, jac=True, method="L-BFGS-B", options=dict(maxiter=10000))
    x, status, nit, nfev = _run_host_lbfgs(lib, fun, u0)
    assert (status in (1, 2)) == bool(r.success)
    assert (nit, nfev) == (r.nit, r.nfev)
    np.testing.assert_allclose(x, r.x, rtol=1e-8, atol=1e-10)


def test_maxiter_reports_failure(lib):
    def ros(x):
        return sopt.rosen(x), sopt.rosen_der(x)
    x0 = np.array([-1.2, 1, 0.5, 2, -1])
    r = sopt.minimize(ros, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=5))
    x, status, nit, nfev = _run_host_lbfgs(lib, ros, x0, maxiter=5)
    assert status == 3 and not r.success
    assert nit == r.nit
    np.testing.assert_allclose(x, r.x, rtol=1e-9)
