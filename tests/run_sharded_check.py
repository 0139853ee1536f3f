"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_check.py
Every rank runs the sharded LocalExpertOI driver; rank 0 then re-runs the same problem on its own GPU
alone and the two sets of tables must be identical (same kernels, same per-expert arithmetic)."""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpsat_b200 import synthetic  # noqa: E402
from gpsat_b200.local_experts import LocalExpertOI  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = synthetic.workload("tiny", n_experts=12)
    df = pd.DataFrame({c: w["table"][i] for i, c in enumerate(w["table_cols"])})
    eloc = pd.DataFrame(w["experts"], columns=w["expert_cols"])
    ploc = pd.DataFrame({c: w["pred"][i] for i, c in enumerate(w["pred_cols"])})
    cfg = dict(expert_loc_config={"source": eloc},
               data_config={"data_source": df, "obs_col": w["obs_col"], "coords_col": w["coords_col"],
                            "local_select": w["local_select"]},
               model_config=w["model"],
               pred_loc_config={"method": "from_dataframe", "df": ploc, "max_dist": w["max_dist"]})
    tabs = LocalExpertOI(device=local, **cfg).run(store_path=None)
    # sparse model: the inducing points are drawn with numpy's global RNG in expert order (gpflow_models.py:809-819);
    # every rank draws for the WHOLE list so a shard sees the single-GPU draws, and the points come back in the gather
    cfg_s = dict(cfg, model_config=dict(w["model"], oi_model="B200SGPRModel",
                                        init_params=dict(w["model"]["init_params"], num_inducing_points=40)))
    np.random.seed(7)
    tabs_s = LocalExpertOI(device=local, **cfg_s).run(store_path=None)
    dist.barrier()
    rank = dist.get_rank()
    world = dist.get_world_size()
    dist.destroy_process_group()
    if rank == 0:
        np.random.seed(7)
        single_s = LocalExpertOI(device=local, **cfg_s).run(store_path=None)
        assert len(tabs_s["inducing_points"]) == len(single_s["inducing_points"]) > 0
        for k in ("run_details", "preds", "lengthscales", "kernel_variance", "likelihood_variance", "inducing_points"):
            a, b = tabs_s[k], single_s[k]
            assert a.index.equals(b.index), "sparse " + k
            for c in a.columns:
                if c != "run_time" and a[c].dtype.kind == "f":
                    np.testing.assert_array_equal(a[c].values, b[c].values, err_msg=f"sparse {k}.{c}")
        print(f"sparse sharded run == single-GPU run: OK ({len(tabs_s['inducing_points'])} inducing-point rows)")
        single = LocalExpertOI(device=local, **cfg).run(store_path=None)
        for k in ("run_details", "preds", "lengthscales", "kernel_variance", "likelihood_variance"):
            a, b = tabs[k], single[k]
            assert a.index.equals(b.index), k
            for c in a.columns:
                if c in ("run_time",):
                    continue
                if a[c].dtype.kind == "f":
                    np.testing.assert_array_equal(a[c].values, b[c].values, err_msg=f"{k}.{c}")
                else:
                    assert (a[c].values == b[c].values).all(), f"{k}.{c}"
        print(f"sharded run over {world} GPUs == single-GPU run: OK ({len(tabs['run_details'])} experts, "
              f"{len(tabs['preds'])} prediction rows)")


if __name__ == "__main__":
    main()
