"""CPU oracle for the GPSat local-expert OI hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy float64 restatement of
the reference's algorithm (GPSat wrappers + the GPflow-2.9 / tfp / scipy
arithmetic they call).  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- never by ``gpsat_b200`` (the product path fails loudly when the
CUDA library is missing; it has no CPU fallback).

Pinning status (see DESIGN.md "Oracle"):
  * GPR LML / predict (rows K1, L1, F1): pinned against sklearn (reference test
    tests/test_localexperts.py:22-49,204-227, KAT-1), the notebook value
    docs/notebooks/gp_regression.ipynb LML=16.6180 (KAT-3) and the reference's
    own PurePythonGPR executed in the authoring container under module stubs
    (tests/golden/make_golden.py).
  * selection (rows S2, S3): pinned against the reference's own
    DataLoader.local_data_select (real scipy KDTree) and
    PredictionLocations._max_dist_bool executed under stubs, plus the
    notebook counts 62/59/41/37/44/38 (KAT-4).
  * optimiser trajectory (row P1): restated from scipy L-BFGS-B; pinned against
    scipy.optimize.minimize itself (same f/g callable).  GPflow's own rounding
    (matmul-form r^2) cannot be executed here: "gpflow form" is restated, not
    pinned, and differs from the direct form at the 1e-7..1e-9 level.
  * post-processing (oracle/postproc.py): pinned against the
    reference's own gaussian_2d_weight, get_weighted_values and
    glue_local_predictions_1d/_2d outputs (tests/golden/postproc.npz).
"""
