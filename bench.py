#!/usr/bin/env python
"""bench.py -- experts/sec (optimise + predict) of the local-expert OI hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU restatement of the reference path)

A "step" is one pass of the hot path (prediction-location filter -> observation selection ->
gather -> L-BFGS optimisation of every expert -> objective -> predictive mean / variance) over one
batch of experts of the named workload.  Experts are independent, so ranks take disjoint batches
(weak scaling, no data-path collective); the only collective is the timing reduction.
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

# per-launch DRAM traffic (bytes) of the dominant kernel of each phase, from the committed ncu --set full captures
# (profiles/*_summary.md); filled in when a capture of the current kernel generation exists
TRAFFIC = {
    # k_potrf_panel, panel J = 4 of 17 (2561 CTAs, 197 slots, N ~ 2.1k) from the ncu --set full capture
    # profiles/r01f_summary.md: 1.64 GB read + 0.31 GB written per launch.  Algorithmic bytes of that launch (every
    # CTA reads its L row panel and the L column panel once, the K tiles once, and writes its L tiles): 2.6 GB, i.e.
    # L2 already absorbs part of the operand re-reads (26 % sector hit rate)
    "potrf": 1.951e9,
    "trtri": 2.595e9,     # k_trtri_pass1, level h = 8 (64 x 197 CTAs), profiles/r01f_busy.md
    "lauum": 3.995e9,     # k_lauum2 (120 x 197 CTAs), profiles/r01f_busy.md
}

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
    ap.add_argument("--cpu-sample", type=int, default=3,
                    help="experts timed by the cpu_baseline leg / per step of --impl reference (~6 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = os.cpu_count() or 1
    return int(n)


def median_experts(w, k):
    """k experts around the middle of the list (representative N)."""
    E = len(w["experts"])
    lo = max(0, E // 2 - k // 2)
    return w["experts"][lo:lo + k]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpsat_b200 import synthetic
    w = synthetic.workload(args.workload)
    k = max(1, args.cpu_sample)
    for _ in range(min(args.warmup, 1)):         # BLAS thread pools / imports
        cpu_experts_per_sec(w, median_experts(w, 1))
    vals, secs = [], 0.0
    for s in range(args.steps):
        lo = (len(w["experts"]) // 2 + s * k) % (len(w["experts"]) - k)
        v, dt, _ = cpu_experts_per_sec(w, w["experts"][lo:lo + k])
        vals.append(k)
        secs += dt
    value = sum(vals) / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['describe']}", "experts_per_step": k},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                             "sample": f"{k} expert(s) per step x {args.steps} steps from the middle of the "
                                       f"{w['name']} expert list; sequential oracle loop (numpy/LAPACK + scipy "
                                       "L-BFGS-B); the reference's GPflow/TensorFlow stack is not installable here"},
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
    from gpsat_b200 import build, get_engine, synthetic
    from gpsat_b200.batched import ModelSpec, run_experts, run_experts_host, h2d_bytes
    build.build()
    eng = get_engine(local)
    dev = eng.device
    w = synthetic.workload(args.workload)
    spec = ModelSpec.from_model_config(w["model"])
    B = min(args.experts_per_step, len(w["experts"]))
    E_all = len(w["experts"])
    n_chunks = max(1, E_all // B)
    # fixed random order: every batch is a statistically identical sample of the lattice, so per-GPU work
    # does not depend on which part of the (density-varying) domain a rank happens to get
    experts_perm = w["experts"][np.random.default_rng(12345).permutation(E_all)]

    perm = np.random.default_rng(12345).permutation(E_all)
    theta_all = w.get("theta")          # predict-only workloads: per-expert hyper-parameters to load

    def chunk(step):
        c = (step * world + rank) % n_chunks
        return np.ascontiguousarray(experts_perm[c * B:(c + 1) * B])

    def chunk_theta(step):
        if theta_all is None:
            return None
        c = (step * world + rank) % n_chunks
        return np.ascontiguousarray(theta_all[perm[c * B:(c + 1) * B]])

    table_h = torch.from_numpy(w["table"]).pin_memory()
    pred_h = torch.from_numpy(w["pred"]).pin_memory()
    table_d, pred_d = table_h.to(dev), pred_h.to(dev)
    kw = dict(table_cols=w["table_cols"], obs_col=w["obs_col"], coords_col=w["coords_col"],
              ref_cols=w["expert_cols"], local_select=w["local_select"], pred_cols=w["pred_cols"],
              max_dist=w["max_dist"], optimise=w["optimise"])

    def step_device(step):
        refs = torch.from_numpy(chunk(step)).to(dev)
        return run_experts(eng, spec, table_d, refs_dev=refs, pred_table_dev=pred_d, theta_init=chunk_theta(step), **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ev0.record()
    n_done, nfev_sum, nobs, nfev_max = 0, 0, [], 0
    for s in range(args.steps):
        r = step_device(args.warmup + s)
        n_done += r["n_valid"]
        nfev_sum += int(r["nfev"].sum().item()) if "nfev" in r else 0
        nfev_max = max(nfev_max, int(r["nfev"].max().item()) if "nfev" in r else 0)
        nobs.append(r["num_obs"].float().mean().item())
    ev1.record()
    barrier()
    ms = maxreduce(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - l0
    total_experts = sumreduce(float(n_done))
    value = total_experts / (ms * 1e-3)

    # ---- one more step with per-phase CUDA events (single stream, so the phases do not overlap) ----
    eng.set_profiling(True)
    pe0, pe1 = torch.cuda.Event(True), torch.cuda.Event(True)
    pe0.record()
    step_device(args.warmup + args.steps - 1)
    pe1.record()
    torch.cuda.synchronize()
    ms_prof = pe0.elapsed_time(pe1)
    prof = eng.get_profile()
    eng.set_profiling(False)

    # ---- end to end through the host-buffer API ----
    refs_h = [torch.from_numpy(chunk(args.warmup + args.steps + s)).pin_memory() for s in range(args.steps)]
    th_h = [chunk_theta(args.warmup + args.steps + s) for s in range(args.steps)]
    run_experts_host(eng, spec, table_h, experts=refs_h[0], pred_table=pred_h, theta_init=th_h[0], **kw)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    n_e2e, d2h = 0, 0
    for s in range(args.steps):
        r = run_experts_host(eng, spec, table_h, experts=refs_h[s], pred_table=pred_h, theta_init=th_h[s], **kw)
        n_e2e += r["n_valid"]
        d2h = sum(v.nbytes for v in r.values() if isinstance(v, np.ndarray))
    ev1.record()
    barrier()
    ms_e2e = maxreduce(ev0.elapsed_time(ev1))
    wall_e2e = maxreduce((time.perf_counter() - t0) * 1e3)
    ms_e2e = max(ms_e2e, wall_e2e)
    e2e_value = sumreduce(float(n_e2e)) / (ms_e2e * 1e-3)
    h2d = h2d_bytes(table_h, pred_h, refs_h[0])

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
    roofline = {"bound": "tensor", "kernel": {"potrf": "k_potrf_panel (batched blocked Cholesky, one launch per panel)",
                                              "trtri": "k_trtri_pass1/2 (triangular inverse)",
                                              "lauum": "k_lauum2 (K^-1 tiles)"}[dom],
                "achieved": phases[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": (phases[dom]["tflops"] / peak) if phases[dom]["tflops"] else None,
                "traffic": TRAFFIC.get(dom),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the phase's dominant kernel, "
                                "from the committed ncu --set full capture (profiles/); null when not captured for "
                                "this kernel generation",
                "peak_source": f"FP64 measured on this GPU in this run: DMMA register-chain {dmma_peak:.1f} TF, "
                               f"cuBLAS DGEMM 4096^3 {dgemm_peak:.1f} TF (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_model": "sum over active experts of N^3/3 per phase per objective evaluation",
                "phases": phases, "share_of_step": {k: v["ms"] / ms_prof for k, v in phases.items()},
                "measured": "CUDA events around the phases of every optimiser round of one extra (untimed) step run "
                            "on a single stream; the timed steps overlap several slot groups on separate streams",
                "cholesky_fp64_tflops": phases["potrf"]["tflops"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        k = max(1, args.cpu_sample)
        v, dt, _ = cpu_experts_per_sec(w, median_experts(w, k))
        cpu = {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"{k} expert(s) from the middle of the {w['name']} expert list, sequential oracle loop "
                         f"(numpy/LAPACK + scipy L-BFGS-B), {dt:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['describe']}", "experts_per_step_per_gpu": B,
                       "batching": "each step = one batched optimise+predict call over a fixed-seed random "
                                   f"sample of {B} of the {E_all} lattice experts per GPU",
                       "mean_obs_per_expert": float(np.mean(nobs)), "mean_nfev": nfev_sum / max(n_done, 1), "max_nfev": nfev_max,
                       "l2": "inputs larger than L2 (factor workspaces are GBs per step)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
