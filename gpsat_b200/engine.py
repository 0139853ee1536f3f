"""Batched local-expert GPR engine: thin Python layer over the C ABI (include/gpsat_b200.h).

torch is used only for device memory, streams and host<->device copies; every numerical
kernel is hand-written CUDA inside libgpsat_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MAXD, MAXP, KERNEL_IDS


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class ExpertBatch:
    """CSR batch of experts resident in device memory (see gpsat_batch in the header)."""
    offsets_host: np.ndarray        # int64 [E+1]
    offsets_dev: torch.Tensor       # int64 [E+1]
    coords_dev: torch.Tensor        # float64 [sumN, D]
    obs_dev: torch.Tensor           # float64 [sumN]
    D: int
    kernel: str = "Matern32"
    coords_scale: Sequence[float] = (1.0,)
    obs_scale: float = 1.0
    obs_mean_local: bool = False
    obs_mean_dev: Optional[torch.Tensor] = None

    @property
    def n_experts(self):
        return len(self.offsets_host) - 1

    def c_struct(self):
        b = _lib.Batch()
        b.n_experts = self.n_experts
        b.D = self.D
        b.kernel_id = KERNEL_IDS[self.kernel]
        b.obs_mean_local = int(self.obs_mean_local)
        b.offsets_host = self.offsets_host.ctypes.data
        b.offsets_dev = self.offsets_dev.data_ptr()
        b.coords_dev = self.coords_dev.data_ptr()
        b.obs_dev = self.obs_dev.data_ptr()
        cs = list(np.broadcast_to(np.asarray(self.coords_scale, dtype=np.float64).ravel(), (self.D,))) \
            if np.size(self.coords_scale) in (1, self.D) else None
        assert cs is not None, "coords_scale must be a scalar or have one entry per coordinate"
        for d in range(MAXD):
            b.coords_scale[d] = float(cs[d]) if d < self.D else 1.0
        b.obs_scale = float(self.obs_scale)
        b.obs_mean_out_dev = self.obs_mean_dev.data_ptr() if self.obs_mean_dev is not None else None
        return b


class Engine:
    """Owns a gpsat_handle on one CUDA device."""

    def __init__(self, device: int = 0, mem_budget_bytes: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("gpsat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        _lib.check(self.lib.gpsat_create(C.byref(h), device, mem_budget_bytes))
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.gpsat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batches ----
    def make_batch(self, offsets, coords, obs, kernel="Matern32", coords_scale=1.0, obs_scale=1.0,
                   obs_mean_local=False, pinned=False) -> ExpertBatch:
        """offsets: int64 [E+1]; coords: [sumN, D]; obs: [sumN]  (numpy or torch, host or device)."""
        offsets_host = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))

        def dev(x, dtype):
            if isinstance(x, torch.Tensor):
                return x.to(self.device, dtype=dtype, non_blocking=True).contiguous()
            return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device, non_blocking=True)

        coords_dev = dev(coords, torch.float64)
        if coords_dev.ndim == 1:
            coords_dev = coords_dev[:, None].contiguous()
        obs_dev = dev(obs, torch.float64).reshape(-1)
        D = coords_dev.shape[1]
        assert D <= MAXD, f"at most {MAXD} coordinate dimensions"
        assert coords_dev.shape[0] == obs_dev.shape[0] == offsets_host[-1]
        E = len(offsets_host) - 1
        return ExpertBatch(offsets_host=offsets_host, offsets_dev=dev(offsets_host, torch.int64),
                           coords_dev=coords_dev, obs_dev=obs_dev, D=D, kernel=kernel,
                           coords_scale=coords_scale, obs_scale=obs_scale, obs_mean_local=obs_mean_local,
                           obs_mean_dev=torch.zeros(E, dtype=torch.float64, device=self.device))

    def make_batch_dev(self, offsets_host, offsets_dev, coords_dev, obs_dev, kernel="Matern32", coords_scale=1.0,
                       obs_scale=1.0, obs_mean_local=False) -> ExpertBatch:
        """Batch over buffers that already live on the device (no copies)."""
        E = len(offsets_host) - 1
        return ExpertBatch(offsets_host=np.ascontiguousarray(offsets_host, dtype=np.int64), offsets_dev=offsets_dev,
                           coords_dev=coords_dev, obs_dev=obs_dev, D=coords_dev.shape[1], kernel=kernel,
                           coords_scale=coords_scale, obs_scale=obs_scale, obs_mean_local=obs_mean_local,
                           obs_mean_dev=torch.zeros(E, dtype=torch.float64, device=self.device))

    def _theta_dev(self, theta, E, D):
        th = torch.as_tensor(np.asarray(theta, dtype=np.float64)) if not isinstance(theta, torch.Tensor) else theta
        th = th.to(self.device, dtype=torch.float64)
        if th.ndim == 1:
            th = th[None, :].expand(E, -1)
        out = torch.zeros(E, MAXP, dtype=torch.float64, device=self.device)
        out[:, :D + 2] = th[:, :D + 2]
        return out.contiguous()

    # ---- L1 / G1 ----
    def eval(self, batch: ExpertBatch, theta, grad=True):
        """-LML [E] and d(-LML)/dtheta [E, D+2] at constrained theta [E, D+2] (or [D+2])."""
        E, D = batch.n_experts, batch.D
        th = self._theta_dev(theta, E, D)
        f = torch.empty(E, dtype=torch.float64, device=self.device)
        g = torch.zeros(E, MAXP, dtype=torch.float64, device=self.device) if grad else None
        b = batch.c_struct()
        _lib.check(self.lib.gpsat_gpr_eval(self.h, C.byref(b), _ptr(th), _ptr(f), _ptr(g), _stream_ptr(self.device)))
        return f, (g[:, :D + 2] if grad else None)

    # ---- P1 ----
    def optimise(self, batch: ExpertBatch, theta0, kind, low, high, trainable, maxiter=10_000, maxfun=15_000,
                 maxcor=10, maxls=20, ftol=2.220446049250313e-09, gtol=1e-5):
        E, D = batch.n_experts, batch.D
        th0 = self._theta_dev(theta0, E, D)
        tr = _lib.Transforms()
        for p in range(D + 2):
            tr.kind[p] = int(kind[p])
            tr.low[p] = float(low[p])
            tr.high[p] = float(high[p])
            tr.trainable[p] = int(bool(trainable[p]))
        oo = _lib.OptOptions(maxcor, maxiter, maxfun, maxls, ftol, gtol)
        theta = torch.zeros(E, MAXP, dtype=torch.float64, device=self.device)
        fobj = torch.empty(E, dtype=torch.float64, device=self.device)
        status = torch.zeros(E, dtype=torch.int32, device=self.device)
        nit = torch.zeros(E, dtype=torch.int32, device=self.device)
        nfev = torch.zeros(E, dtype=torch.int32, device=self.device)
        b = batch.c_struct()
        _lib.check(self.lib.gpsat_gpr_optimise(self.h, C.byref(b), _ptr(th0), C.byref(tr), C.byref(oo), _ptr(theta),
                                               _ptr(fobj), _ptr(status), _ptr(nit), _ptr(nfev),
                                               _stream_ptr(self.device)))
        return {"theta": theta[:, :D + 2], "theta_full": theta, "fobj": fobj, "status": status, "nit": nit,
                "nfev": nfev}

    # ---- F1 ----
    def predict(self, batch: ExpertBatch, theta, pred_offsets, pred_coords, pred_offsets_dev=None):
        E, D = batch.n_experts, batch.D
        th = self._theta_dev(theta, E, D)
        poff_host = np.ascontiguousarray(np.asarray(pred_offsets, dtype=np.int64))
        assert len(poff_host) == E + 1
        poff_dev = torch.as_tensor(poff_host).to(self.device) if pred_offsets_dev is None else pred_offsets_dev
        if isinstance(pred_coords, torch.Tensor):
            pc = pred_coords.to(self.device, dtype=torch.float64).contiguous()
            if pc.ndim == 1:
                pc = pc[:, None].contiguous()
        else:
            pc = torch.as_tensor(np.ascontiguousarray(pred_coords, dtype=np.float64)).to(self.device)
        P = int(poff_host[-1])
        fmean = torch.empty(P, dtype=torch.float64, device=self.device)
        fvar = torch.empty(P, dtype=torch.float64, device=self.device)
        yvar = torch.empty(P, dtype=torch.float64, device=self.device)
        fobj = torch.empty(E, dtype=torch.float64, device=self.device)
        b = batch.c_struct()
        _lib.check(self.lib.gpsat_gpr_predict(self.h, C.byref(b), _ptr(th), C.c_void_p(poff_host.ctypes.data),
                                              _ptr(poff_dev), _ptr(pc), _ptr(fmean), _ptr(fvar), _ptr(yvar),
                                              _ptr(fobj), _stream_ptr(self.device)))
        return fmean, fvar, yvar, fobj

    def predict_full_cov(self, batch: ExpertBatch, theta, pred_coords):
        """posterior mean [P] and full covariance [P, P] of f* for the first expert of the batch"""
        D = batch.D
        th = self._theta_dev(theta, batch.n_experts, D)
        pc = torch.as_tensor(np.ascontiguousarray(pred_coords, dtype=np.float64)).to(self.device)
        if pc.ndim == 1:
            pc = pc[:, None].contiguous()
        P = pc.shape[0]
        fmean = torch.empty(P, dtype=torch.float64, device=self.device)
        fcov = torch.empty(P, P, dtype=torch.float64, device=self.device)
        b = batch.c_struct()
        _lib.check(self.lib.gpsat_gpr_predict_cov(self.h, C.byref(b), _ptr(th), _ptr(pc), P, _ptr(fmean), _ptr(fcov),
                                                  _stream_ptr(self.device)))
        return fmean, fcov

    # ---- SG1: sparse GPR ----
    def make_sgpr_batch(self, batch: ExpertBatch, z_offsets, z_coords):
        """Attach inducing points (CSR, raw coordinates) to a data batch."""
        zoff_h = np.ascontiguousarray(np.asarray(z_offsets, dtype=np.int64))
        assert len(zoff_h) == batch.n_experts + 1
        zc = z_coords if isinstance(z_coords, torch.Tensor) else torch.as_tensor(
            np.ascontiguousarray(z_coords, dtype=np.float64))
        zc = zc.to(self.device, dtype=torch.float64).contiguous()
        if zc.ndim == 1:
            zc = zc[:, None].contiguous()
        assert zc.shape == (int(zoff_h[-1]), batch.D)
        return {"batch": batch, "zoff_host": zoff_h, "zoff_dev": torch.as_tensor(zoff_h).to(self.device),
                "zcoords": zc}

    @staticmethod
    def _sgpr_struct(sb):
        s = _lib.SgprBatch()
        s.data = sb["batch"].c_struct()
        s.z_offsets_host = sb["zoff_host"].ctypes.data
        s.z_offsets_dev = sb["zoff_dev"].data_ptr()
        s.z_coords_dev = sb["zcoords"].data_ptr()
        return s

    def sgpr_eval(self, sb, theta, grad=True):
        """-ELBO [E] and d(-ELBO)/dtheta [E, D+2]."""
        batch = sb["batch"]
        E, D = batch.n_experts, batch.D
        th = self._theta_dev(theta, E, D)
        f = torch.empty(E, dtype=torch.float64, device=self.device)
        g = torch.zeros(E, MAXP, dtype=torch.float64, device=self.device) if grad else None
        s = self._sgpr_struct(sb)
        _lib.check(self.lib.gpsat_sgpr_eval(self.h, C.byref(s), _ptr(th), _ptr(f), _ptr(g), _stream_ptr(self.device)))
        return f, (g[:, :D + 2] if grad else None)

    def sgpr_optimise(self, sb, theta0, kind, low, high, trainable, maxiter=10_000, maxfun=15_000, maxcor=10,
                      maxls=20, ftol=2.220446049250313e-09, gtol=1e-5):
        batch = sb["batch"]
        E, D = batch.n_experts, batch.D
        th0 = self._theta_dev(theta0, E, D)
        tr = _lib.Transforms()
        for p in range(D + 2):
            tr.kind[p] = int(kind[p])
            tr.low[p] = float(low[p])
            tr.high[p] = float(high[p])
            tr.trainable[p] = int(bool(trainable[p]))
        oo = _lib.OptOptions(maxcor, maxiter, maxfun, maxls, ftol, gtol)
        theta = torch.zeros(E, MAXP, dtype=torch.float64, device=self.device)
        fobj = torch.empty(E, dtype=torch.float64, device=self.device)
        status = torch.zeros(E, dtype=torch.int32, device=self.device)
        nit = torch.zeros(E, dtype=torch.int32, device=self.device)
        nfev = torch.zeros(E, dtype=torch.int32, device=self.device)
        s = self._sgpr_struct(sb)
        _lib.check(self.lib.gpsat_sgpr_optimise(self.h, C.byref(s), _ptr(th0), C.byref(tr), C.byref(oo), _ptr(theta),
                                                _ptr(fobj), _ptr(status), _ptr(nit), _ptr(nfev),
                                                _stream_ptr(self.device)))
        return {"theta": theta[:, :D + 2], "theta_full": theta, "fobj": fobj, "status": status, "nit": nit,
                "nfev": nfev}

    def sgpr_predict(self, sb, theta, pred_offsets, pred_coords):
        batch = sb["batch"]
        E, D = batch.n_experts, batch.D
        th = self._theta_dev(theta, E, D)
        poff_host = np.ascontiguousarray(np.asarray(pred_offsets, dtype=np.int64))
        assert len(poff_host) == E + 1
        poff_dev = torch.as_tensor(poff_host).to(self.device)
        pc = pred_coords if isinstance(pred_coords, torch.Tensor) else torch.as_tensor(
            np.ascontiguousarray(pred_coords, dtype=np.float64))
        pc = pc.to(self.device, dtype=torch.float64).contiguous()
        if pc.ndim == 1:
            pc = pc[:, None].contiguous()
        P = int(poff_host[-1])
        fmean = torch.empty(P, dtype=torch.float64, device=self.device)
        fvar = torch.empty(P, dtype=torch.float64, device=self.device)
        yvar = torch.empty(P, dtype=torch.float64, device=self.device)
        fobj = torch.empty(E, dtype=torch.float64, device=self.device)
        s = self._sgpr_struct(sb)
        _lib.check(self.lib.gpsat_sgpr_predict(self.h, C.byref(s), _ptr(th), C.c_void_p(poff_host.ctypes.data),
                                               _ptr(poff_dev), _ptr(pc), _ptr(fmean), _ptr(fvar), _ptr(yvar),
                                               _ptr(fobj), _stream_ptr(self.device)))
        return fmean, fvar, yvar, fobj

    # ---- K1 ----
    def kernel_matrix(self, X1, X2, theta, kernel="Matern32", coords_scale=None, add_noise=False):
        x1 = torch.as_tensor(np.ascontiguousarray(X1, dtype=np.float64)).to(self.device)
        x2 = torch.as_tensor(np.ascontiguousarray(X2, dtype=np.float64)).to(self.device)
        if coords_scale is not None:
            cs = torch.as_tensor(np.asarray(coords_scale, dtype=np.float64)).to(self.device)
            x1 = (x1 / cs).contiguous()
            x2 = (x2 / cs).contiguous()
        D = x1.shape[1]
        th = torch.zeros(MAXP, dtype=torch.float64, device=self.device)
        th[:D + 2] = torch.as_tensor(np.asarray(theta, dtype=np.float64)).to(self.device)
        K = torch.empty(x1.shape[0], x2.shape[0], dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gpsat_kernel_matrix(_ptr(x1), x1.shape[0], _ptr(x2), x2.shape[0], D,
                                                KERNEL_IDS[kernel], _ptr(th), int(add_noise), _ptr(K),
                                                _stream_ptr(self.device)))
        return K

    def debug_factor(self, batch: ExpertBatch, theta):
        """Dense (nb*64)^2 L_aug and X = L_aug^-1 of expert 0 (test hook)."""
        D = batch.D
        n0 = int(batch.offsets_host[1] - batch.offsets_host[0])
        npad = (n0 // 64 + 1) * 64
        th = self._theta_dev(theta, batch.n_experts, D)
        L = torch.empty(npad, npad, dtype=torch.float64, device=self.device)
        X = torch.empty(npad, npad, dtype=torch.float64, device=self.device)
        b = batch.c_struct()
        _lib.check(self.lib.gpsat_debug_factor(self.h, C.byref(b), _ptr(th), _ptr(L), _ptr(X),
                                               _stream_ptr(self.device)))
        return L, X

    # ---- S2 / S3 ----
    def select_count(self, spec: "_lib.SelSpec", table_dev: torch.Tensor, refs_dev: torch.Tensor):
        """matches per expert (int64 [E]); table_dev [ncols, n] column-major, refs_dev [E, nrefcols]."""
        assert table_dev.is_contiguous() and refs_dev.is_contiguous()
        E, nref = refs_dev.shape
        counts = torch.zeros(E, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.gpsat_select_count(C.byref(spec), _ptr(table_dev), table_dev.shape[1], _ptr(refs_dev),
                                               nref, E, _ptr(counts), _stream_ptr(self.device)))
        return counts

    def select_fill(self, spec: "_lib.SelSpec", table_dev: torch.Tensor, refs_dev: torch.Tensor,
                    offsets: torch.Tensor, total: int):
        """matching row indices (int32 [total]) in ascending order per expert at offsets[e]."""
        assert table_dev.is_contiguous() and refs_dev.is_contiguous()
        E, nref = refs_dev.shape
        idx = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.gpsat_select_fill(C.byref(spec), _ptr(table_dev), table_dev.shape[1], _ptr(refs_dev),
                                              nref, E, _ptr(offsets), _ptr(idx), _stream_ptr(self.device)))
        return idx[:total]

    # ---- grid-bucketed S2 / S3 ----
    def build_buckets(self, spec: "_lib.SelSpec", table_dev: torch.Tensor):
        """Bin the table rows on the spec's two-column ball / max_dist term (None if it has none)."""
        term = next((spec.t[k] for k in range(spec.nterms) if spec.t[k].type in (1, 2) and spec.t[k].ncol == 2), None)
        n = table_dev.shape[1]
        if term is None or n == 0 or not (term.val > 0):
            return None
        xs, ys = table_dev[term.col[0]], table_dev[term.col[1]]
        lo = torch.stack([xs.min(), ys.min(), xs.max(), ys.max()]).cpu().numpy()
        if not np.isfinite(lo).all():
            return None
        cell = float(term.val) * 1.000001
        ncx = int((lo[2] - lo[0]) / cell) + 1
        ncy = int((lo[3] - lo[1]) / cell) + 1
        if ncx * ncy > (1 << 26):
            return None
        g = _lib.CellGrid(float(lo[0]), float(lo[1]), cell, ncx, ncy)
        st = _stream_ptr(self.device)
        counts = torch.zeros(ncx * ncy, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.gpsat_bucket_build(C.byref(spec), C.byref(g), _ptr(table_dev), n, 0, _ptr(counts), None,
                                               None, st))
        start = torch.zeros(ncx * ncy + 1, dtype=torch.int64, device=self.device)
        start[1:] = torch.cumsum(counts, 0)
        counts.zero_()
        order = torch.empty(n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.gpsat_bucket_build(C.byref(spec), C.byref(g), _ptr(table_dev), n, 1, _ptr(counts),
                                               _ptr(start), _ptr(order), st))
        return {"grid": g, "start": start, "order": order}

    def select_count_bucketed(self, spec, buckets, table_dev, refs_dev):
        E, nref = refs_dev.shape
        counts = torch.zeros(E, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.gpsat_select_bucket(C.byref(spec), C.byref(buckets["grid"]), _ptr(table_dev),
                                                table_dev.shape[1], _ptr(refs_dev), nref, E, _ptr(buckets["start"]),
                                                _ptr(buckets["order"]), 0, _ptr(counts), None, None,
                                                _stream_ptr(self.device)))
        return counts

    def select_fill_bucketed(self, spec, buckets, table_dev, refs_dev, offsets, total, max_count):
        E, nref = refs_dev.shape
        idx = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.gpsat_select_bucket(C.byref(spec), C.byref(buckets["grid"]), _ptr(table_dev),
                                                table_dev.shape[1], _ptr(refs_dev), nref, E, _ptr(buckets["start"]),
                                                _ptr(buckets["order"]), int(max_count), None, _ptr(offsets),
                                                _ptr(idx), _stream_ptr(self.device)))
        return idx[:total]

    def select(self, spec: "_lib.SelSpec", table_dev: torch.Tensor, refs_dev: torch.Tensor):
        """Returns (offsets int64 [E+1], idx int32 [total]) on the device; indices ascending per expert."""
        counts = self.select_count(spec, table_dev, refs_dev)
        offsets = torch.zeros(refs_dev.shape[0] + 1, dtype=torch.int64, device=self.device)
        offsets[1:] = torch.cumsum(counts, 0)
        total = int(offsets[-1].item())
        return offsets, self.select_fill(spec, table_dev, refs_dev, offsets, total)

    def gather_rows(self, table_dev: torch.Tensor, idx: torch.Tensor, coord_cols, obs_col):
        """coords [total, D] (row-major) and obs [total] of the selected rows of a [ncols, n] table."""
        total, D = int(idx.numel()), len(coord_cols)
        coords = torch.empty(total, D, dtype=torch.float64, device=self.device)
        obs = torch.empty(total, dtype=torch.float64, device=self.device)
        cc = (C.c_int * D)(*[int(c) for c in coord_cols])
        _lib.check(self.lib.gpsat_gather_rows(_ptr(table_dev), table_dev.shape[1], _ptr(idx), total, D, cc,
                                              int(obs_col), _ptr(coords), _ptr(obs), _stream_ptr(self.device)))
        return coords, obs

    def gather_pred(self, table_dev: torch.Tensor, refs_dev: torch.Tensor, offsets: torch.Tensor,
                    idx: torch.Tensor, table_cols, ref_cols):
        """Prediction coords [total, D]: table column table_cols[d] (>= 0) or the expert's refs[:, ref_cols[d]]."""
        total, D = int(idx.numel()), len(table_cols)
        out = torch.empty(total, D, dtype=torch.float64, device=self.device)
        tc = (C.c_int * D)(*[int(c) for c in table_cols])
        rc = (C.c_int * D)(*[int(c) for c in ref_cols])
        _lib.check(self.lib.gpsat_gather_pred(_ptr(table_dev), table_dev.shape[1], _ptr(refs_dev),
                                              refs_dev.shape[1], refs_dev.shape[0], _ptr(offsets), _ptr(idx), D,
                                              tc, rc, _ptr(out), _stream_ptr(self.device)))
        return out

    # ---- counters ----
    def dmma_peak_tflops(self, iters=20000):
        """FP64 DMMA issue-rate speed of light of this GPU (TFLOP/s), see gpsat_dmma_peak."""
        tf, ms = C.c_double(), C.c_double()
        _lib.check(self.lib.gpsat_dmma_peak(self.device.index, iters, C.byref(tf), C.byref(ms)))
        return tf.value

    def launch_count(self):
        return int(self.lib.gpsat_launch_count(self.h))

    def sync_timeouts(self):
        return int(self.lib.gpsat_sync_timeouts(self.h))

    def last_plan(self):
        """slot plan of the last batched call: resident experts, tile rows of the largest matrix, memory"""
        s, nb, per, bud = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        _lib.check(self.lib.gpsat_last_plan(self.h, C.byref(s), C.byref(nb), C.byref(per), C.byref(bud)))
        return {"slots": s.value, "max_obs_padded": nb.value * 64, "bytes_per_slot": per.value,
                "workspace_bytes": s.value * per.value, "budget_bytes": bud.value}

    def set_profiling(self, on=True):
        _lib.check(self.lib.gpsat_set_profiling(self.h, int(on)))

    def get_profile(self):
        v = [C.c_double() for _ in range(9)]
        _lib.check(self.lib.gpsat_get_profile(self.h, *[C.byref(x) for x in v]))
        k = ["ms_potrf", "ms_trtri", "ms_lauum", "ms_other", "flops_potrf", "flops_trtri", "flops_lauum",
             "ms_build", "ms_trace"]
        return {a: b.value for a, b in zip(k, v)}


def make_sel_spec(terms):
    """terms: list of dicts {type, cols, rcols, comp, val}."""
    comp_code = {">=": 0, ">": 1, "==": 2, "<": 3, "<=": 4}
    sp = _lib.SelSpec()
    assert len(terms) <= _lib.SEL_MAXTERMS
    sp.nterms = len(terms)
    for k, t in enumerate(terms):
        st = sp.t[k]
        st.type = t["type"]
        st.ncol = len(t["cols"])
        st.comp = comp_code.get(t.get("comp", "<="), 4)
        for j, (c, rc) in enumerate(zip(t["cols"], t["rcols"])):
            st.col[j] = c
            st.rcol[j] = rc
        st.val = float(t["val"])
    return sp
