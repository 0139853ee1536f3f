"""Multi-GPU sharding of the expert list (one process per GPU, torch.distributed).

Experts are independent: the observation and prediction tables are replicated, every rank
computes the per-expert observation / prediction-location counts redundantly (two bucketed count kernels), derives the
SAME longest-processing-time partition by N^3 cost without communicating, runs its shard through ``run_experts`` with
everything resident on its GPU, and the per-expert results are gathered ONCE: each rank packs all of its result arrays
into one device buffer of 8-byte words and a single ``all_gather`` (NCCL over NVLink on the GPUs, gloo in the CPU
tests) moves the payload -- preceded only by the 4-word header exchange that sizes it.  Nothing on the data path
leaves the device before that gather; the gathered payload is copied to the host once and merged into global expert
order there (the next consumer is pandas).
"""
from __future__ import annotations

import time
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


# ---------------------------------------------------------------------------------------------
# the packed result payload
# ---------------------------------------------------------------------------------------------
# (name, per-what, trailing width: "D" / "D+2" / int, dtype); per-what in E (shard experts), V (valid experts),
# P (prediction rows), Z (inducing-point rows).  Every field travels as 8-byte words.
FIELDS = (("num_obs", "E", 1, np.int64), ("has_pred", "E", 1, np.int64), ("too_few", "E", 1, np.int64),
          ("valid", "E", 1, np.int64), ("valid_idx", "V", 1, np.int64), ("theta", "V", "D+2", np.float64),
          ("fobj", "V", 1, np.float64), ("obs_mean", "V", 1, np.float64), ("status", "V", 1, np.int64),
          ("nit", "V", 1, np.int64), ("nfev", "V", 1, np.int64), ("pred_count", "V", 1, np.int64),
          ("z_count", "V", 1, np.int64), ("pred_coords", "P", "D", np.float64), ("fmean", "P", 1, np.float64),
          ("fvar", "P", 1, np.float64), ("yvar", "P", 1, np.float64), ("inducing_points", "Z", "D", np.float64))
PER_EXPERT = ("theta", "fobj", "status", "nit", "nfev", "obs_mean")
PER_PRED = ("pred_coords", "fmean", "fvar", "yvar")
_INT32 = ("status", "nit", "nfev")


def _width(w, D):
    return {"D": D, "D+2": D + 2}.get(w, w)


def pack_results(res: Dict[str, torch.Tensor], D: int, device) -> (torch.Tensor, torch.Tensor):
    """One shard's results -> (header int64[4] = E, V, P, Z; payload int64[words]) on ``device``.
    Fields the run did not produce (no optimisation / prediction / sparse model) travel as zeros."""
    E = int(res["num_obs"].shape[0])
    V = int(res.get("n_valid", 0))
    live = V > 0 and "theta" in res
    P = int(res["pred_coords"].shape[0]) if live and "pred_coords" in res else 0
    Z = int(res["inducing_points"].shape[0]) if live and "inducing_points" in res else 0
    n = {"E": E, "V": V if live else 0, "P": P, "Z": Z}
    extra = {}
    if live and "pred_offsets" in res:
        extra["pred_count"] = torch.diff(torch.as_tensor(res["pred_offsets"]))
    if live and "z_offsets" in res:
        extra["z_count"] = torch.diff(torch.as_tensor(res["z_offsets"]))
    words = []
    for name, per, w, dt in FIELDS:
        rows, width = n[per], _width(w, D)
        t = extra.get(name, res.get(name) if (per == "E" or live) else None)
        if t is None or rows == 0:
            words.append(torch.zeros(rows * width, dtype=torch.int64, device=device))
            continue
        t = torch.as_tensor(t).to(device)
        t = t.to(torch.float64) if dt is np.float64 else t.to(torch.int64)
        assert t.numel() == rows * width, (name, tuple(t.shape), rows, width)
        words.append(t.contiguous().view(torch.int64).reshape(-1))
    header = torch.tensor([E, n["V"], P, Z], dtype=torch.int64, device=device)
    return header, torch.cat(words) if words else torch.zeros(0, dtype=torch.int64, device=device)


def unpack_results(header: np.ndarray, payload: np.ndarray, D: int) -> Dict[str, np.ndarray]:
    """Inverse of pack_results on host arrays (header int64[4], payload int64[>= words])."""
    E, V, P, Z = (int(x) for x in header)
    n = {"E": E, "V": V, "P": P, "Z": Z}
    out, o = {}, 0
    for name, per, w, dt in FIELDS:
        rows, width = n[per], _width(w, D)
        a = payload[o:o + rows * width].view(dt)
        o += rows * width
        out[name] = a.reshape(rows, width) if width != 1 or isinstance(w, str) else a
    for k in ("has_pred", "too_few", "valid"):
        out[k] = out[k].astype(bool)
    for k in _INT32:
        out[k] = out[k].astype(np.int32)
    out["n_valid"] = V
    out["pred_offsets"] = np.concatenate([[0], np.cumsum(out.pop("pred_count"))]).astype(np.int64)
    out["z_offsets"] = np.concatenate([[0], np.cumsum(out.pop("z_count"))]).astype(np.int64)
    return out


def gather_results(res: Dict[str, torch.Tensor], D: int, device, group=None) -> List[Dict[str, np.ndarray]]:
    """The only data collective of a run: every rank's packed results on every rank (list indexed by rank)."""
    world = dist.get_world_size(group)
    header, payload = pack_results(res, D, device)
    headers = torch.empty(world, 4, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(headers, header[None, :].contiguous(), group=group)
    hh = headers.cpu().numpy()
    sizes = [sum(int(h[{"E": 0, "V": 1, "P": 2, "Z": 3}[per]]) * _width(w, D) for _, per, w, _ in FIELDS) for h in hh]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=torch.int64, device=device)
    pad[:payload.numel()] = payload
    allp = torch.empty(world, m, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allp, pad[None, :], group=group)
    host = allp.cpu().numpy()                      # one D2H copy of the whole gathered payload
    return [unpack_results(hh[r], host[r], D) for r in range(world)]


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

    def ragged(offs_key, keys):
        """re-order CSR rows of the concatenated shards into global expert order"""
        cnt = np.concatenate([np.diff(p[offs_key]) for p in live])
        base = np.cumsum([0] + [int(p[offs_key][-1]) for p in live])[:-1]
        starts = np.concatenate([np.asarray(p[offs_key][:-1]) + b for p, b in zip(live, base)])
        cnt_o, starts_o = cnt[order], starts[order]
        off = np.zeros(len(cnt_o) + 1, dtype=np.int64)
        off[1:] = np.cumsum(cnt_o)
        # row r of the merged array comes from starts_o[e] + (r - off[e]) for the expert e that owns it
        take = np.repeat(np.asarray(starts_o, dtype=np.int64) - off[:-1], cnt_o) + np.arange(off[-1], dtype=np.int64)
        out[offs_key] = off
        for k in keys:
            out[k] = np.concatenate([p[k] for p in live])[take]

    if all("pred_offsets" in p for p in live):
        ragged("pred_offsets", PER_PRED)
    if all("z_offsets" in p and p["z_offsets"][-1] > 0 for p in live):
        ragged("z_offsets", ("inducing_points",))
    return out


# timing of the last sharded call on this rank (bench.py reports load imbalance and the gather cost from it)
LAST = {}


def run_experts_sharded(eng, spec, table, table_cols, obs_col, coords_col, experts, ref_cols, local_select,
                        group=None, **kw) -> Optional[dict]:
    """``run_experts_host`` over all ranks of ``group``; returns the merged result on every rank."""
    from .batched import inducing_local_rows, run_experts, run_experts_host, sel_terms
    from .engine import make_sel_spec
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    experts = np.ascontiguousarray(experts, dtype=np.float64)
    if world == 1:
        return run_experts_host(eng, spec, table, table_cols, obs_col, coords_col, experts, ref_cols, local_select,
                                **kw)
    dev = eng.device
    D = len(coords_col)

    def up(x):
        if x is None or isinstance(x, torch.Tensor):
            return None if x is None else x.to(dev)
        return torch.from_numpy(np.require(x, dtype=np.float64, requirements=["C", "W"])).to(dev)

    tab_d, refs_d, pred_d = up(table), up(experts), up(kw.pop("pred_table", None))
    pred_cols, max_dist = kw.get("pred_cols"), kw.get("max_dist")
    min_obs = kw.get("min_obs", 3)
    # ---- cost model: observations per expert (every rank computes the same counts) ----
    ospec = make_sel_spec(sel_terms(local_select, list(table_cols), list(ref_cols)))
    obk = eng.build_buckets(ospec, tab_d)
    ocount = (eng.select_count_bucketed(ospec, obk, tab_d, refs_d) if obk else
              eng.select_count(ospec, tab_d, refs_d)).cpu().numpy()
    shards = partition_lpt(ocount.astype(np.float64) ** 3, world)
    mine = shards[rank]
    theta_init = kw.pop("theta_init", None)
    if theta_init is not None:
        theta_init = np.asarray(theta_init)[mine]
    # ---- sparse model: the inducing-point draws of the WHOLE list, in list order (same as one GPU) ----
    inducing_local = None
    if spec.num_inducing_points is not None and not kw.get("count_only", False):
        if pred_d is not None and max_dist is not None:
            found = [c for c in coords_col if c in pred_cols]
            pspec = make_sel_spec([{"type": 2, "cols": [pred_cols.index(c) for c in found],
                                    "rcols": [ref_cols.index(c) for c in found], "val": max_dist}])
            pbk = eng.build_buckets(pspec, pred_d)
            pcount = (eng.select_count_bucketed(pspec, pbk, pred_d, refs_d) if pbk else
                      eng.select_count(pspec, pred_d, refs_d)).cpu().numpy()
        else:
            pcount = np.full(len(experts), 1 if pred_d is None else pred_d.shape[1])
        valid = (pcount > 0) & (ocount >= min_obs)
        draws = inducing_local_rows(ocount[valid], spec.num_inducing_points)
        where = np.cumsum(valid) - 1
        inducing_local = [draws[where[i]] for i in mine if valid[i]]
    t0 = time.perf_counter()
    res = run_experts(eng, spec, tab_d, table_cols, obs_col, coords_col, refs_d[torch.as_tensor(mine, device=dev)],
                      ref_cols, local_select, pred_table_dev=pred_d, theta_init=theta_init,
                      inducing_local=inducing_local, **kw)
    torch.cuda.synchronize(dev) if dev.type == "cuda" else None
    t1 = time.perf_counter()
    # ---- the only data collective: one packed payload per rank ----
    parts = gather_results(res, D, dev, group)
    t2 = time.perf_counter()
    LAST.update(shard_experts=len(mine), shard_cost=float((ocount[mine].astype(np.float64) ** 3).sum()),
                total_cost=float((ocount.astype(np.float64) ** 3).sum()), compute_s=t1 - t0, gather_s=t2 - t1,
                payload_words=int(sum(len(p["num_obs"]) for p in parts)))
    out = merge_shards(shards, parts, len(experts))
    if kw.get("count_only", False) or out["n_valid"] == 0:
        return out
    if not kw.get("optimise", True):
        for k in _INT32:
            out.pop(k, None)
    if not kw.get("predict", True):
        for k in PER_PRED + ("pred_offsets",):
            out.pop(k, None)
    if spec.num_inducing_points is None:
        out.pop("inducing_points", None)
        out.pop("z_offsets", None)
    return out
