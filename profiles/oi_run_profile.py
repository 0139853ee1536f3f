"""Wall time and host profile of the user-facing call, LocalExpertOI.run, on the small-matrix workloads (run via
gpurun): c2 = predict-only with loaded hyper-parameters (363 experts, ~1.7 M prediction rows), c1 = optimise + predict
(256 experts).  Prints the tables' row counts, the wall time of a warm run and the top of a cProfile of it."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import synthetic  # noqa: E402
from gpsat_b200.local_experts import LocalExpertOI  # noqa: E402


def frames(w):
    df = pd.DataFrame({c: w["table"][i] for i, c in enumerate(w["table_cols"])})
    eloc = pd.DataFrame(np.asarray(w["experts"]), columns=w["expert_cols"])
    ploc = pd.DataFrame({c: w["pred"][i] for i, c in enumerate(w["pred_cols"])})
    return df, eloc, ploc


def param_tables(w, eloc):
    cc, th = w["coords_col"], w["theta"]
    D = len(cc)
    idx = pd.MultiIndex.from_frame(eloc[cc])
    rep = pd.MultiIndex.from_frame(eloc[cc].loc[eloc.index.repeat(D)])
    return {"lengthscales": pd.DataFrame({"_dim_0": np.tile(np.arange(D), len(eloc)), "lengthscales": th[:, :D].ravel()},
                                         index=rep),
            "kernel_variance": pd.DataFrame({"_dim_0": 0, "kernel_variance": th[:, D]}, index=idx),
            "likelihood_variance": pd.DataFrame({"_dim_0": 0, "likelihood_variance": th[:, D + 1]}, index=idx)}


def run(name, optimise):
    w = synthetic.workload(name)
    df, eloc, ploc = frames(w)
    model = dict(w["model"])
    if not optimise:
        model["load_params"] = {"file": param_tables(w, eloc)}
    oi = LocalExpertOI(expert_loc_config={"source": eloc},
                       data_config={"data_source": df, "obs_col": w["obs_col"], "coords_col": w["coords_col"],
                                    "local_select": w["local_select"]},
                       model_config=model,
                       pred_loc_config={"method": "from_dataframe", "df": ploc, "max_dist": w["max_dist"]})
    oi.run(store_path=None, optimise=optimise)          # warm-up (library load, buffers)
    t0 = time.perf_counter()
    tabs = oi.run(store_path=None, optimise=optimise)
    dt = time.perf_counter() - t0
    print(f"{name}: LocalExpertOI.run {dt * 1e3:.1f} ms for {len(eloc)} experts = {len(eloc) / dt:.0f} experts/s; rows: "
          + ", ".join(f"{k} {len(v)}" for k, v in tabs.items()))
    pr = cProfile.Profile()
    pr.enable()
    oi.run(store_path=None, optimise=optimise)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
    print("\n".join(ln for ln in s.getvalue().splitlines() if ln.strip())[:6000])


if __name__ == "__main__":
    run("c2", False)
    run("c1", True)
