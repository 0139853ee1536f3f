"""Seeded synthetic workloads of the shapes BASELINE.json names.

The reference's example data blobs are not distributed (.MISSING_LARGE_BLOBS), so every benchmark
and full-size parity input is synthetic: observations binned on a 50 km EASE2-like lattice inside
the lat >= 60 disk (cf. examples/inline_example.py:176-183), integer-day ``t`` around 18326
(2020-03-05) with the +-4 day window of configs/example_local_expert_oi.json:68-87, and
``obs`` = smooth field + N(0, 0.05^2).
"""
from __future__ import annotations

import numpy as np

SEED = 20200305
T0 = 18326.0
CELL = 50_000.0

LOCAL_SELECT = [{"col": "t", "comp": "<=", "val": 4}, {"col": "t", "comp": ">=", "val": -4},
                {"col": ["x", "y"], "comp": "<", "val": 300_000}]

MODEL_C3 = {"oi_model": "B200GPRModel",
            "init_params": {"coords_scale": [50_000, 50_000, 1]},
            "constraints": {"lengthscales": {"low": [1e-8, 1e-8, 1e-8], "high": [600_000, 600_000, 9]}}}
MODEL_C1 = {"oi_model": "B200GPRModel",
            "init_params": {"coords_scale": [50_000, 50_000, 1]},
            "constraints": {"lengthscales": {"low": [1e-8, 1e-8, 1e-8], "high": [600_000, 600_000, 9]},
                            "likelihood_variance": {"low": 0.00125, "high": 0.01}}}


def _field(rng, x, y, t):
    """sum of 3 random plane waves (amplitude 0.1 m overall), slowly varying in time"""
    f = np.zeros_like(x)
    for _ in range(3):
        k = rng.normal(0, 1.0 / 400_000.0, 2)
        w = rng.normal(0, 1.0 / 15.0)
        ph = rng.uniform(0, 2 * np.pi)
        f += (0.1 / np.sqrt(3)) * np.sin(k[0] * x + k[1] * y + w * (t - T0) + ph) * np.sqrt(2)
    return f


def observations(rng, radius_cells, days, density_lo, density_hi, n_sat=3, jitter=20_000.0):
    """One row per (cell, day, satellite) kept with a smooth spatially varying probability."""
    g = np.arange(-radius_cells, radius_cells + 1)
    cx, cy = np.meshgrid(g, g, indexing="ij")
    keep = cx * cx + cy * cy <= radius_cells * radius_cells
    cx, cy = cx[keep].astype(np.float64) * CELL, cy[keep].astype(np.float64) * CELL
    # smooth density field in [density_lo, density_hi]
    ang, ph = rng.uniform(0, 2 * np.pi, 2)
    kx, ky = np.cos(ang) / 0.9e6, np.sin(ang) / 0.9e6
    dens = 0.5 * (density_lo + density_hi) + 0.5 * (density_hi - density_lo) * np.sin(kx * cx + ky * cy + ph)
    xs, ys, ts = [], [], []
    for d in days:
        for _ in range(n_sat):
            m = rng.random(len(cx)) < dens
            n = int(m.sum())
            xs.append(cx[m] + rng.uniform(-jitter, jitter, n))
            ys.append(cy[m] + rng.uniform(-jitter, jitter, n))
            ts.append(np.full(n, float(d)))
    x, y, t = np.concatenate(xs), np.concatenate(ys), np.concatenate(ts)
    obs = _field(rng, x, y, t) + rng.normal(0, 0.05, len(x))
    perm = rng.permutation(len(x))           # source order is not spatially sorted
    return np.ascontiguousarray(np.stack([x[perm], y[perm], t[perm], obs[perm]]))


def expert_lattice(n_experts, spacing=CELL, t=T0):
    """the n_experts lattice points nearest the pole, in raster order (y, then x) like a sorted csv"""
    r = int(np.ceil(np.sqrt(n_experts / np.pi))) + 2
    g = np.arange(-r, r + 1)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    d2 = (gx * gx + gy * gy).ravel()
    order = np.argsort(d2, kind="stable")[:n_experts]
    order.sort()
    ex, ey = gx.ravel()[order] * spacing, gy.ravel()[order] * spacing
    return np.ascontiguousarray(np.column_stack([ex, ey, np.full(n_experts, t)]).astype(np.float64))


def pred_grid(half_width, spacing=5_000.0):
    g = np.arange(-half_width, half_width + spacing / 2, spacing)
    px, py = np.meshgrid(g, g, indexing="ij")
    return np.ascontiguousarray(np.stack([px.ravel(), py.ravel()]).astype(np.float64))


def workload(name="c3", n_experts=None, seed=SEED):
    """dict(table [4, n] rows x,y,t,obs; experts [E, 3]; pred [2, n_pred]; configs ...)."""
    rng = np.random.default_rng(seed)
    if name == "c3":
        # pan-Arctic 50 km expert grid, 300 km radius / 9-day window, N ~ 1000-2000, full optimisation
        E = 8192 if n_experts is None else n_experts
        table = observations(rng, radius_cells=60, days=range(18316, 18337), density_lo=0.31, density_hi=0.68)
        experts = expert_lattice(E)
        half = float(np.abs(experts[:, :2]).max()) + 30_000.0
        return dict(name="c3", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half), pred_cols=["x", "y"], max_dist=30_000.0,
                    local_select=LOCAL_SELECT, model=MODEL_C3, coords_col=["x", "y", "t"], obs_col="obs",
                    optimise=True,
                    describe="pan-Arctic 50 km expert lattice, 300 km radius / 9-day window, ~1-2k obs per "
                             "expert, Matern32 ARD (x,y,t), L-BFGS optimise + predict on the 5 km grid within 30 km")
    if name == "c1":
        # inline_example-like: 200 km expert lattice, N ~ 400-600, P ~ 5027 (5 km grid within 200 km)
        E = 256 if n_experts is None else n_experts
        table = observations(rng, radius_cells=40, days=range(18316, 18337), density_lo=0.14, density_hi=0.22)
        experts = expert_lattice(E, spacing=200_000.0)
        half = float(np.abs(experts[:, :2]).max()) + 200_000.0
        return dict(name="c1", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half), pred_cols=["x", "y"], max_dist=200_000.0,
                    local_select=LOCAL_SELECT, model=MODEL_C1, coords_col=["x", "y", "t"], obs_col="obs",
                    optimise=True,
                    describe="inline_example shape: 200 km expert lattice, ~400-600 obs per expert, "
                             "optimise + predict on the 5 km grid within 200 km")
    if name == "c2":
        # configs/example_local_expert_oi.json shape: 363 expert locations (the reference's
        # example_expert_locations_arctic_no_date.csv has 363 rows), N ~ 400-600, fixed per-expert (smoothed)
        # hyper-parameters loaded instead of optimised, prediction on the 5 km grid within 200 km (P ~ 5027)
        E = 363 if n_experts is None else n_experts
        table = observations(rng, radius_cells=40, days=range(18316, 18337), density_lo=0.14, density_hi=0.22)
        experts = expert_lattice(E, spacing=200_000.0)
        half = float(np.abs(experts[:, :2]).max()) + 200_000.0
        # KAT-5-like ranges of the smoothed tables: l ~ (5, 3, 9), kernel variance ~ 0.015, noise ~ 0.003
        theta = np.column_stack([rng.uniform(3.0, 8.0, E), rng.uniform(2.0, 6.0, E), rng.uniform(5.0, 9.0, E),
                                 rng.uniform(0.008, 0.03, E), rng.uniform(0.002, 0.006, E)])
        return dict(name="c2", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half), pred_cols=["x", "y"], max_dist=200_000.0,
                    local_select=LOCAL_SELECT, model=MODEL_C1, coords_col=["x", "y", "t"], obs_col="obs",
                    optimise=False, theta=theta,
                    describe="example_local_expert_oi.json shape: 363 experts, ~400-600 obs per expert, fixed "
                             "(loaded) hyper-parameters, predict-only on the 5 km grid within 200 km (P ~ 5027)")
    if name == "c4":
        # 5 km-resolution expert grid (the full run has ~1e5 experts; a step takes a lattice sample of them), dense
        # along-track observations: variable N up to ~8k per expert, full optimisation
        E = 4096 if n_experts is None else n_experts
        table = observations(rng, radius_cells=60, days=range(18316, 18337), density_lo=0.25, density_hi=0.98,
                             n_sat=8)
        experts = expert_lattice(E)            # every 10th point of the 5 km grid: the density range is sampled
        half = float(np.abs(experts[:, :2]).max()) + 10_000.0
        return dict(name="c4", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half, 2_500.0), pred_cols=["x", "y"], max_dist=5_000.0,
                    local_select=LOCAL_SELECT, model=MODEL_C3, coords_col=["x", "y", "t"], obs_col="obs",
                    optimise=True,
                    describe="5 km expert grid, 300 km radius / 9-day window, dense tracks: variable N up to ~8k obs "
                             "per expert, Matern32 ARD (x,y,t), L-BFGS optimise + predict within 5 km")
    if name == "c5":
        # sparse path: GPflowSGPRModel-equivalent, 500 inducing points per expert, N ~ 2k-10k, 50 km expert lattice
        E = 8192 if n_experts is None else n_experts
        table = observations(rng, radius_cells=60, days=range(18316, 18337), density_lo=0.18, density_hi=0.85,
                             n_sat=12)
        experts = expert_lattice(E)
        half = float(np.abs(experts[:, :2]).max()) + 30_000.0
        model = dict(MODEL_C3, oi_model="B200SGPRModel",
                     init_params=dict(MODEL_C3["init_params"], num_inducing_points=500))
        return dict(name="c5", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half), pred_cols=["x", "y"], max_dist=30_000.0,
                    local_select=LOCAL_SELECT, model=model, coords_col=["x", "y", "t"], obs_col="obs",
                    optimise=True,
                    describe="sparse GPR, 500 inducing points per expert, 300 km radius / 9-day window, "
                             "~2-10k obs per expert, L-BFGS optimise + predict on the 5 km grid within 30 km")
    if name == "tiny":
        E = 16 if n_experts is None else n_experts
        table = observations(rng, radius_cells=14, days=range(18322, 18331), density_lo=0.10, density_hi=0.16)
        experts = expert_lattice(E, spacing=100_000.0)
        half = float(np.abs(experts[:, :2]).max()) + 50_000.0
        return dict(name="tiny", table=table, table_cols=["x", "y", "t", "obs"], experts=experts,
                    expert_cols=["x", "y", "t"], pred=pred_grid(half, 10_000.0), pred_cols=["x", "y"],
                    max_dist=50_000.0, local_select=LOCAL_SELECT, model=MODEL_C1, coords_col=["x", "y", "t"],
                    obs_col="obs", optimise=True, describe="smoke-sized workload")
    raise ValueError(name)
