#!/usr/bin/env python
"""bench.py -- experts/sec (optimise + predict) of the local-expert OI hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU restatement of the reference path)

A "step" is one pass of the hot path (prediction-location filter -> observation selection ->
gather -> L-BFGS optimisation of every expert -> objective -> predictive mean / variance) over one
LIST of experts of the named workload.  With N GPUs the list holds N x --experts-per-step experts and goes through
gpsat_b200.distributed.run_experts_sharded: every rank derives the same LPT partition by N^3 cost, runs its shard,
and ONE packed all_gather returns the results to every rank (weak scaling: per-GPU work is fixed as N grows; at
N = 8 the default list is the whole 8192-expert lattice).  --strong fixes the list at --total-experts instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# per-launch DRAM traffic (bytes) of each phase's dominant kernel comes from profiles/r02_traffic.json, written by
# profiles/capture_traffic.sh from an `ncu --set full` capture of THIS command line's batch configuration; when the
# file is missing or was captured for another experts-per-step the key is null (a number from another batch size
# says nothing about this run)
def load_traffic(experts_per_step, workload):
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("experts_per_step") == experts_per_step and t.get("workload") == workload:
            return t
    except (OSError, ValueError):
        pass
    return None


METRIC = "experts/sec (optimise+predict)"
UNIT = "experts/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c1", "c2", "c4", "c5", "tiny"])
    ap.add_argument("--experts-per-step", type=int, default=1024, help="experts per rank per step")
    ap.add_argument("--cpu-sample", type=int, default=32,
                    help="experts of the fixed CPU sample (the first K of the fixed-seed expert order): timed by the "
                         "cpu_baseline leg, and spread over the steps of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="sweeps only: leave the host-buffer leg out (e2e: null)")
    ap.add_argument("--strong", action="store_true", help="one fixed list of --total-experts for every N")
    ap.add_argument("--total-experts", type=int, default=8192)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's sequential loop (restated reference; gpflow/tensorflow are not installable)
# ------------------------------------------------------------------------------------------------
def cpu_experts_per_sec(w, expert_rows, n_threads_note=True):
    import pandas as pd
    from oracle.local_expert_oi import run_local_expert_oi
    df = pd.DataFrame({c: w["table"][i] for i, c in enumerate(w["table_cols"])})
    eloc = pd.DataFrame(expert_rows, columns=w["expert_cols"])
    ploc = pd.DataFrame({c: w["pred"][i] for i, c in enumerate(w["pred_cols"])})
    data = {"data_source": df, "obs_col": w["obs_col"], "coords_col": w["coords_col"],
            "local_select": w["local_select"]}
    kw = {}
    model = {k: v for k, v in w["model"].items() if k != "oi_model"}
    if str(w["model"].get("oi_model", "")).endswith("SGPRModel"):
        from oracle.sgpr import OracleSGPRModel
        kw["model_cls"] = OracleSGPRModel
    t0 = time.perf_counter()
    _, per = run_local_expert_oi(eloc, data, model, {"method": "from_dataframe", "df": ploc,
                                                     "max_dist": w["max_dist"]}, optimise=w["optimise"], **kw)
    dt = time.perf_counter() - t0
    return len(per) / dt, dt, per


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_cores():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must not be measured on one thread"""
    n = cpu_cores()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = cpu_cores()
    return int(n)


def expert_order(w):
    """the fixed-seed order both arms draw their experts from (batches are statistically identical samples of the
    density-varying lattice; the CPU sample is its first K entries)"""
    return np.random.default_rng(12345).permutation(len(w["experts"]))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpsat_b200 import synthetic
    w = synthetic.workload(args.workload)
    cores = use_all_cores()
    order = expert_order(w)
    k = max(1, -(-args.cpu_sample // max(args.steps, 1)))       # ceil: the steps together cover >= cpu_sample experts
    for _ in range(min(args.warmup, 1)):         # BLAS thread pools / imports
        cpu_experts_per_sec(w, w["experts"][order[-1:]])
    n, secs = 0, 0.0
    for s in range(args.steps):
        idx = order[(s * k) % len(order):(s * k) % len(order) + k]
        v, dt, per = cpu_experts_per_sec(w, w["experts"][idx])
        n += len(per)
        secs += dt
    value = n / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['describe']}", "experts_per_step": k},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads": cpu_threads(), "kind": "port",
                             "sample": f"the first {n} experts of the fixed-seed {w['name']} expert order ({k} per step "
                                       f"x {args.steps} steps); sequential oracle loop (numpy/LAPACK + scipy "
                                       "L-BFGS-B) on all host cores; the reference's GPflow/TensorFlow stack is not "
                                       "installable here"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gpsat_b200 import build, distributed, get_engine, synthetic
    from gpsat_b200.batched import ModelSpec, run_experts, run_experts_host, h2d_bytes
    build.build()
    eng = get_engine(local)
    dev = eng.device
    w = synthetic.workload(args.workload)
    spec = ModelSpec.from_model_config(w["model"])
    E_all = len(w["experts"])
    order = expert_order(w)
    if args.strong:
        L = min(args.total_experts, E_all)              # one fixed list for every N
        B = -(-L // world)
    else:
        B = min(args.experts_per_step, E_all)           # per GPU
        L = min(B * world, E_all)
    theta_all = w.get("theta")          # predict-only workloads: per-expert hyper-parameters to load

    def list_idx(step):
        """the step's expert list: the next L entries of the fixed-seed order (wrapping around the lattice)"""
        lo = (step * L) % E_all
        return np.r_[order[lo:lo + L], order[:max(0, lo + L - E_all)]]

    table_h = torch.from_numpy(w["table"]).pin_memory()
    pred_h = torch.from_numpy(w["pred"]).pin_memory()
    table_d, pred_d = table_h.to(dev), pred_h.to(dev)
    kw = dict(table_cols=w["table_cols"], obs_col=w["obs_col"], coords_col=w["coords_col"],
              ref_cols=w["expert_cols"], local_select=w["local_select"], pred_cols=w["pred_cols"],
              max_dist=w["max_dist"], optimise=w["optimise"])
    shard_stats = []

    def step_device(step):
        """inputs resident in HBM: the observation / prediction tables stay on the device, only the step's expert
        rows (KBs) are new.  N > 1: LPT shard + the single packed result gather are inside the step."""
        idx = list_idx(step)
        th = None if theta_all is None else np.ascontiguousarray(theta_all[idx])
        if world == 1:
            refs = torch.from_numpy(np.ascontiguousarray(w["experts"][idx])).to(dev)
            return run_experts(eng, spec, table_d, refs_dev=refs, pred_table_dev=pred_d, theta_init=th, **kw)
        r = distributed.run_experts_sharded(eng, spec, table_d, experts=w["experts"][idx], pred_table=pred_d,
                                            theta_init=th, **kw)
        shard_stats.append(dict(distributed.LAST))
        return r

    def step_host(step):
        """end to end through the host-buffer API: tables from pinned host memory, results back as numpy"""
        idx = list_idx(step)
        th = None if theta_all is None else np.ascontiguousarray(theta_all[idx])
        if world == 1:
            return run_experts_host(eng, spec, table_h, experts=w["experts"][idx], pred_table=pred_h, theta_init=th,
                                    **kw)
        return distributed.run_experts_sharded(eng, spec, table_h, experts=w["experts"][idx], pred_table=pred_h,
                                               theta_init=th, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    maxreduce = lambda x: reduce(x, dist.ReduceOp.MAX)
    sumreduce = lambda x: reduce(x, dist.ReduceOp.SUM)

    def scalar(v):
        return float(v.float().mean().item()) if isinstance(v, torch.Tensor) else float(np.mean(v))

    # FP64 roofline denominators measured on this GPU before the run
    dmma_peak = eng.dmma_peak_tflops()
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    torch.matmul(a, a)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(3):
        e0.record()
        torch.matmul(a, a)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dgemm_peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
    del a

    # ---- device-resident timing ----
    for s in range(args.warmup):
        step_device(s)
    shard_stats.clear()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    ev0.record()
    n_done, nfev_sum, nobs, nfev_max = 0, 0, [], 0
    for s in range(args.steps):
        r = step_device(args.warmup + s)
        n_done += int(r["n_valid"])
        if "nfev" in r:
            nf = r["nfev"]
            nfev_sum += int(nf.sum())
            nfev_max = max(nfev_max, int(nf.max()))
        nobs.append(scalar(r["num_obs"]))
    ev1.record()
    slot_plan = eng.last_plan()
    barrier()
    # device events on this rank's stream, max over ranks; the sharded step ends with host work (merge), so the
    # wall clock is taken too and the larger of the two counts
    ms = max(maxreduce(ev0.elapsed_time(ev1)), maxreduce((time.perf_counter() - t0) * 1e3) if world > 1 else 0.0)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - l0
    total_experts = float(n_done) if world > 1 else float(n_done)     # sharded results are already global
    value = total_experts / (ms * 1e-3)
    sharding = None
    if world > 1:
        comp = np.array([s_["compute_s"] for s_ in shard_stats])
        gath = np.array([s_["gather_s"] for s_ in shard_stats])
        cost = shard_stats[-1]["shard_cost"]
        sharding = {"mode": ("strong: one fixed list of %d experts" % L) if args.strong else
                            ("one list of %d x %d experts per step" % (world, B)),
                    "partition": "LPT greedy on N^3 (identical on every rank, no communication)",
                    "collective": "one packed all_gather of the per-expert / per-prediction results per step (NCCL)",
                    "shard_compute_ms_per_step_max": maxreduce(float(comp.mean() * 1e3)),
                    "shard_compute_ms_per_step_mean": sumreduce(float(comp.mean() * 1e3)) / world,
                    "gather_and_wait_ms_per_step_rank0": float(gath.mean() * 1e3),
                    "gather_ms_per_step_min_over_ranks": -maxreduce(-float(gath.mean() * 1e3)),
                    "cost_imbalance_max_over_mean": maxreduce(cost) / (sumreduce(cost) / world)}
        sharding["time_imbalance_max_over_mean"] = (sharding["shard_compute_ms_per_step_max"] /
                                                    sharding["shard_compute_ms_per_step_mean"])

    # ---- one more step with per-phase CUDA events (single stream, so the phases do not overlap) ----
    eng.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(True), torch.cuda.Event(True)
    pe0.record()
    idx = list_idx(args.warmup + args.steps - 1)[rank::world][:min(B, 1024)]
    run_experts(eng, spec, table_d, refs_dev=torch.from_numpy(np.ascontiguousarray(w["experts"][idx])).to(dev),
                pred_table_dev=pred_d, theta_init=None if theta_all is None else np.ascontiguousarray(theta_all[idx]),
                **kw)
    pe1.record()
    torch.cuda.synchronize()
    ms_prof = pe0.elapsed_time(pe1)
    prof = eng.get_profile()
    eng.set_profiling(False)

    # ---- end to end through the host-buffer API ----
    e2e = None
    if not args.skip_e2e:
        step_host(args.warmup + args.steps)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        n_e2e, d2h = 0, 0
        for s in range(args.steps):
            r = step_host(args.warmup + args.steps + 1 + s)
            n_e2e += int(r["n_valid"])
            d2h = sum(v.nbytes for v in r.values() if isinstance(v, np.ndarray))
        ev1.record()
        barrier()
        ms_e2e = max(maxreduce(ev0.elapsed_time(ev1)), maxreduce((time.perf_counter() - t0) * 1e3))
        h2d = h2d_bytes(table_h, pred_h) + L * w["experts"].shape[1] * 8
        e2e = {"value": float(n_e2e) / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel group: the batched Cholesky (k_potrf_panel launches of every round) ----
    phases = {}
    for nm in ("potrf", "trtri", "lauum"):
        t_ms, fl = prof[f"ms_{nm}"], prof[f"flops_{nm}"]
        phases[nm] = {"ms": t_ms, "tflops": (fl / (t_ms * 1e-3) / 1e12) if t_ms > 0 else None}
    for nm in ("build", "trace", "other"):      # FP64-pipe elementwise kernels (no N^3 work): time only
        phases[nm] = {"ms": prof[f"ms_{nm}"], "tflops": None}
    dom = max(("potrf", "trtri", "lauum"), key=lambda k: phases[k]["ms"])
    peak = max(dmma_peak, dgemm_peak)
    traffic = load_traffic(B, w["name"])
    roofline = {"bound": "tensor", "kernel": {"potrf": "k_potrf_panel (batched blocked Cholesky, one launch per panel)",
                                              "trtri": "k_trtri_pass1/2 (triangular inverse)",
                                              "lauum": "k_lauum2 (K^-1 tiles)"}[dom],
                "achieved": phases[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": (phases[dom]["tflops"] / peak) if phases[dom]["tflops"] else None,
                "traffic": None if traffic is None else traffic.get(dom, {}).get("dram_bytes_per_launch"),
                "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the phase's dominant kernel, "
                                 "averaged over every launch of it in one round, from the ncu captures of this batch "
                                 "configuration (profiles/capture_r02.sh -> profiles/r02_traffic.json; the --set full "
                                 "figures of three of the launches are under full_set_capture): " +
                                 json.dumps(traffic.get(dom))) if traffic else
                                "null: no ncu --set full capture of this batch configuration is committed "
                                "(profiles/r02_traffic.json)",
                "peak_source": f"FP64 measured on this GPU in this run: DMMA register-chain {dmma_peak:.1f} TF, "
                               f"cuBLAS DGEMM 4096^3 {dgemm_peak:.1f} TF (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_model": "sum over active experts of N^3/3 per phase per objective evaluation",
                "phases": phases, "share_of_step": {k: v["ms"] / ms_prof for k, v in phases.items()},
                "measured": "CUDA events around the phases of every optimiser round of one extra (untimed) step run "
                            "on a single stream; the timed steps overlap several slot groups on separate streams",
                "cholesky_fp64_tflops": phases["potrf"]["tflops"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # the same code path, sample and thread count as the --impl reference arm, in its own process (the OpenBLAS
        # pool of this process shares the cores with torch's threads and the clock sampler)
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
               "--steps", "1", "--warmup", "1", "--cpu-sample", str(args.cpu_sample)]
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        out = subprocess.run(cmd, capture_output=True, text=True, env=env)
        try:
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception:
            cpu = {"value": None, "unit": UNIT, "cores": cpu_cores(), "kind": "port",
                   "sample": "cpu_baseline leg failed: " + out.stderr[-300:]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['describe']}", "experts_per_step_per_gpu": B,
                       "experts_per_step": L,
                       "batching": f"each step = ONE list of {L} experts (fixed-seed order over the {E_all} lattice "
                                   "experts) through one batched optimise+predict call per GPU" +
                                   ("" if world == 1 else "; the list is LPT-sharded over the ranks and gathered once"),
                       "mean_obs_per_expert": float(np.mean(nobs)), "mean_nfev": nfev_sum / max(n_done, 1),
                       "max_nfev": nfev_max, "sharding": sharding, "slot_plan_rank0": slot_plan,
                       "l2": "inputs larger than L2 (factor workspaces are GBs per step)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
