"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: LPT partition, the packed single-payload gather
(pack -> all_gather -> unpack) and the merge back into global expert order.  No GPU and no engine: every rank fabricates its shard's results from a
deterministic function of the global expert index, the merged result must equal the single-process one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpsat_b200.distributed import gather_results, merge_shards, pack_results, partition_lpt, unpack_results


def test_partition_lpt_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    n = rng.integers(200, 8000, 500).astype(float)
    cost = n ** 3
    for world in (1, 2, 4, 8):
        sh = partition_lpt(cost, world)
        assert sorted(np.concatenate(sh).tolist()) == list(range(500))
        loads = np.array([cost[s].sum() for s in sh])
        assert loads.max() / loads.mean() < 1.02
        assert all((np.diff(s) > 0).all() for s in sh)
        sh2 = partition_lpt(cost.copy(), world)
        assert all((a == b).all() for a, b in zip(sh, sh2))


def _fake_result(gidx, D=3):
    """what run_experts_host would return for the experts `gidx` (global indices)"""
    gidx = np.asarray(gidx, dtype=np.int64)
    num_obs = (gidx * 7) % 50
    has_pred = (gidx % 11) != 0
    too_few = has_pred & (num_obs < 3)
    valid = has_pred & ~too_few
    vidx = np.flatnonzero(valid)
    g = gidx[vidx]
    cnt = 1 + (g % 4)
    poff = np.zeros(len(g) + 1, dtype=np.int64)
    poff[1:] = np.cumsum(cnt)
    rep = np.repeat(g, cnt)
    within = np.concatenate([np.arange(c) for c in cnt]) if len(cnt) else np.zeros(0, dtype=np.int64)
    zcnt = 2 + (g % 3)
    zoff = np.zeros(len(g) + 1, dtype=np.int64)
    zoff[1:] = np.cumsum(zcnt)
    zrep = np.repeat(g, zcnt)
    return dict(z_offsets=zoff, inducing_points=np.stack([zrep * 1.5, zrep - 1.0, zrep * 0.25], axis=1),
                num_obs=num_obs, has_pred=has_pred, too_few=too_few, valid=valid, valid_idx=vidx, n_valid=len(vidx),
                theta=np.stack([g + 0.1 * k for k in range(D + 2)], axis=1).astype(float), fobj=-1.0 * g,
                obs_mean=0.5 * g, status=(g % 3).astype(np.int32), nit=(g % 17).astype(np.int32),
                nfev=(g % 19).astype(np.int32), pred_offsets=poff,
                pred_coords=np.stack([rep + 0.0, within + 0.0, rep * 2.0], axis=1), fmean=rep + 0.25 * within,
                fvar=rep * 3.0 + within, yvar=rep * 5.0 + within)


def _worker(rank, world, port, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cost = ((np.arange(E) * 37) % 101 + 1.0) ** 3
        shards = partition_lpt(cost, world)
        res = _fake_result(shards[rank])
        # the engine hands back device tensors; here they are CPU tensors and the collective runs over gloo
        rt = {k: (torch.as_tensor(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in res.items()}
        parts = gather_results(rt, 3, torch.device("cpu"))
        merged = merge_shards(shards, parts, E)
        if rank == 0:
            q.put({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in merged.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E", [37, 5])
def test_sharded_gather_matches_single_process(E):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, E, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_result(np.arange(E))
    for k, v in ref.items():
        np.testing.assert_array_equal(np.asarray(merged[k]), np.asarray(v), err_msg=k)


def test_pack_unpack_round_trip_including_empty_shards():
    for gidx in ([], [0], [0, 11, 22], list(range(3, 40))):       # multiples of 11 have no prediction locations
        res = _fake_result(np.asarray(gidx, dtype=np.int64))
        rt = {k: (torch.as_tensor(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in res.items()}
        header, payload = pack_results(rt, 3, torch.device("cpu"))
        back = unpack_results(header.numpy(), payload.numpy(), 3)
        for k, v in res.items():
            if res["n_valid"] == 0 and k not in ("num_obs", "has_pred", "too_few", "valid", "n_valid", "valid_idx"):
                continue
            np.testing.assert_array_equal(np.asarray(back[k]).reshape(np.asarray(v).shape), np.asarray(v), err_msg=k)


def _driver_worker(rank, world, port, tmp, q):
    """LocalExpertOI.run under torch.distributed (gloo): rank 0 alone reads / writes the store, the others follow its
    config id and resume decisions.  The engine call is the oracle-backed stand-in of tests/refrun_common.py."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # two processes share this container's cores: BLAS pools (numpy's, and scipy's, which loads later) spin when
    # oversubscribed, so both are capped before and after the imports
    os.environ["OPENBLAS_NUM_THREADS"] = os.environ["OMP_NUM_THREADS"] = "2"
    from threadpoolctl import threadpool_limits
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpsat_b200
        import refrun_common as rc
        import scipy.linalg  # noqa: F401
        from gpsat_b200 import local_experts as le
        threadpool_limits(limits=2)
        rc.fh.install()
        le.run_experts_sharded = rc.oracle_backed_sharded
        gpsat_b200.get_engine = lambda device=0: None
        le.LocalExpertOI._device_name = lambda self: "oracle-cpu"
        cfg, _, store_path = rc.setup_files(os.path.join(tmp, f"rank{rank}"))
        oi = rc.make_oi(le.LocalExpertOI, cfg)
        oi.expert_locs = oi.expert_locs.iloc[[0, 1, 5]].copy(True)       # two ordinary experts + the one without data
        tabs = oi.run(store_path=store_path, return_tables=True, **dict(cfg["run_kwargs"], max_batch=8))
        wrote = sorted(rc.fh.tables(store_path)) if os.path.abspath(store_path) in rc.fh.FILES else []
        q.put((rank, sorted(tabs), len(tabs["run_details"]), wrote))
    finally:
        dist.destroy_process_group()


def test_driver_run_under_two_ranks_only_rank0_touches_the_store(tmp_path):
    for r in range(2):
        os.makedirs(tmp_path / f"rank{r}")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r[0], r[1:]) for r in (q.get(timeout=300), q.get(timeout=300)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks shaped the same tables (3 locations: 2 ran + 1 with too few observations) ...
    assert got[0][0] == got[1][0] and got[0][1] == got[1][1] == 3
    # ... but only rank 0 wrote them (rank 1's store holds nothing but the input table it was seeded with)
    assert "run_details" in got[0][2] and "oi_config" in got[0][2] and "expert_locs" in got[0][2]
    assert got[1][2] == []
