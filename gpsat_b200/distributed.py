"""Multi-GPU sharding of the expert list (one process per GPU, torch.distributed).

Experts are independent: the observation and prediction tables are replicated,
every rank computes the per-expert observation counts redundantly (cheap), derives the SAME
longest-processing-time partition by N^3 cost without communicating, runs its shard through
``run_experts`` and the per-expert results are gathered once at the end (NCCL over NVLink on the
GPUs, gloo in the CPU tests).  There is no collective on the data path.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist


def partition_lpt(cost: np.ndarray, world: int) -> List[np.ndarray]:
    """Greedy longest-processing-time partition; every shard is returned in ascending index order
    so the sequential loop's expert order is preserved inside a shard.  Deterministic."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    shards: List[list] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))          # ties -> lowest rank: identical on every rank
        shards[r].append(int(i))
        load[r] += cost[i]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def _gather_padded(x: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all_gather of first-dimension-ragged tensors (same trailing shape, same dtype)."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    pad = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return [o[:s] for o, s in zip(out, sizes)]


PER_EXPERT = ("theta", "fobj", "status", "nit", "nfev", "obs_mean")
PER_PRED = ("pred_coords", "fmean", "fvar", "yvar")


def merge_shards(shard_idx: List[np.ndarray], parts: List[Dict[str, np.ndarray]], n_experts: int):
    """Combine per-rank results (each over the VALID experts of its shard, ascending global index)
    into one result in global expert order -- the layout ``run_experts_host`` returns."""
    out: Dict[str, np.ndarray] = {}
    num_obs = np.zeros(n_experts, dtype=np.int64)
    has_pred = np.zeros(n_experts, dtype=bool)
    too_few = np.zeros(n_experts, dtype=bool)
    valid = np.zeros(n_experts, dtype=bool)
    gidx = []
    for idx, p in zip(shard_idx, parts):
        num_obs[idx] = p["num_obs"]
        has_pred[idx] = p["has_pred"]
        too_few[idx] = p["too_few"]
        valid[idx] = p["valid"]
        gidx.append(idx[np.asarray(p["valid_idx"], dtype=np.int64)])
    gidx = np.concatenate(gidx) if gidx else np.zeros(0, dtype=np.int64)
    order = np.argsort(gidx, kind="stable")
    out.update(num_obs=num_obs, has_pred=has_pred, too_few=too_few, valid=valid,
               valid_idx=np.flatnonzero(valid), n_valid=int(valid.sum()))
    if len(gidx) == 0:
        return out
    live = [p for p in parts if p["n_valid"]]
    for k in PER_EXPERT:
        if all(k in p for p in live):
            out[k] = np.concatenate([p[k] for p in live])[order]
    if all("pred_offsets" in p for p in live):
        cnt = np.concatenate([np.diff(p["pred_offsets"]) for p in live])
        base = np.cumsum([0] + [int(p["pred_offsets"][-1]) for p in live])[:-1]
        starts = np.concatenate([np.asarray(p["pred_offsets"][:-1]) + b for p, b in zip(live, base)])
        cnt_o, starts_o = cnt[order], starts[order]
        take = np.concatenate([np.arange(s, s + c) for s, c in zip(starts_o, cnt_o)])
        poff = np.zeros(len(cnt_o) + 1, dtype=np.int64)
        poff[1:] = np.cumsum(cnt_o)
        out["pred_offsets"] = poff
        for k in PER_PRED:
            out[k] = np.concatenate([p[k] for p in live])[take]
    return out


def run_experts_sharded(eng, spec, table, table_cols, obs_col, coords_col, experts, ref_cols, local_select,
                        group=None, **kw) -> Optional[dict]:
    """``run_experts_host`` over all ranks of ``group``; returns the merged result on every rank."""
    from .batched import run_experts_host, sel_terms
    from .engine import make_sel_spec
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    experts = np.ascontiguousarray(experts, dtype=np.float64)
    if world == 1:
        return run_experts_host(eng, spec, table, table_cols, obs_col, coords_col, experts, ref_cols, local_select,
                                **kw)
    dev = eng.device
    tab_d = torch.as_tensor(np.ascontiguousarray(table, dtype=np.float64)).to(dev)
    refs_d = torch.as_tensor(experts).to(dev)
    ospec = make_sel_spec(sel_terms(local_select, list(table_cols), list(ref_cols)))
    counts = eng.select_count(ospec, tab_d, refs_d).cpu().numpy().astype(np.float64)
    shards = partition_lpt(counts ** 3, world)
    mine = shards[rank]
    theta_init = kw.pop("theta_init", None)
    if theta_init is not None:
        theta_init = np.asarray(theta_init)[mine]
    res = run_experts_host(eng, spec, table, table_cols, obs_col, coords_col, experts[mine], ref_cols, local_select,
                           theta_init=theta_init, **kw)
    # ---- the only collective: gather the per-expert / per-prediction results ----
    D = len(coords_col)
    empty = {"theta": np.zeros((0, D + 2)), "fobj": np.zeros(0), "obs_mean": np.zeros(0),
             "status": np.zeros(0, dtype=np.int32), "nit": np.zeros(0, dtype=np.int32),
             "nfev": np.zeros(0, dtype=np.int32), "pred_offsets": np.zeros(1, dtype=np.int64),
             "pred_coords": np.zeros((0, D)), "fmean": np.zeros(0), "fvar": np.zeros(0), "yvar": np.zeros(0)}
    keys = ["num_obs", "has_pred", "too_few", "valid", "valid_idx", "theta", "fobj", "obs_mean"]
    if kw.get("optimise", True):
        keys += ["status", "nit", "nfev"]
    if kw.get("predict", True):
        keys += ["pred_offsets"] + list(PER_PRED)
    gathered = {}
    for k in keys:
        v = res[k] if k in res else empty[k]
        t = torch.as_tensor(np.ascontiguousarray(v)).to(dev)
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        gathered[k] = [g.cpu().numpy() for g in _gather_padded(t, group)]
    parts = []
    for r in range(world):
        p = {k: gathered[k][r] for k in keys}
        for k in ("has_pred", "too_few", "valid"):
            p[k] = p[k].astype(bool)
        p["n_valid"] = int(p["valid"].sum())
        parts.append(p)
    return merge_shards(shards, parts, len(experts))
