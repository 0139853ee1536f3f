// gpsat_b200: slot management, parameter transforms and the device-side optimiser step.
//
// A "slot" holds one resident expert (coordinates, observations, factor tiles, optimiser state).
// Experts stream through a fixed pool of slots: when an expert's L-BFGS terminates its slot is
// refilled on the device from a work queue (atomic counter), so the batch stays full while
// iteration counts differ between experts.
#pragma once
#include "gpr2.cuh"
#include "lbfgs.cuh"

namespace gpsat {

// Parameter transforms (gpflow positive() = softplus, likelihood variance =
// softplus + 1e-6, constrained parameters = tfp.bijectors.Sigmoid(low, high))
struct TransformSpec {
  int kind[MAXP];      // 0: theta = softplus(u) + low ; 1: theta = low + (high-low)*sigmoid(u)
  double low[MAXP];
  double high[MAXP];
  int free_idx[MAXP];  // indices of trainable parameters
  int nfree;
  int np;              // D + 2
};

GPSAT_HD inline double sigmoid_d(double u) {
  if (u >= 0.0) return 1.0 / (1.0 + exp(-u));
  const double e = exp(u);
  return e / (1.0 + e);
}
GPSAT_HD inline double tr_fwd(const TransformSpec& t, int p, double u) {
  if (t.kind[p] == 0) return fmax(u, 0.0) + log1p(exp(-fabs(u))) + t.low[p];
  return t.low[p] + (t.high[p] - t.low[p]) * sigmoid_d(u);
}
GPSAT_HD inline double tr_inv(const TransformSpec& t, int p, double th) {
  if (t.kind[p] == 0) {
    const double y = th - t.low[p];
    return y + log(-expm1(-y));
  }
  const double x = (th - t.low[p]) / (t.high[p] - t.low[p]);
  return log(x) - log1p(-x);
}
GPSAT_HD inline double tr_dfwd(const TransformSpec& t, int p, double u) {
  const double s = sigmoid_d(u);
  if (t.kind[p] == 0) return s;
  return (t.high[p] - t.low[p]) * s * (1.0 - s);
}

struct BatchIn {
  int E, D;
  const double* coords;        // [sumN][D] raw coordinates, row-major
  const double* obs;           // [sumN]
  const long long* offsets;    // [E+1]
  const int* order;            // [E] processing order (expert ids) or nullptr
  const double* theta0;        // [E][MAXP] constrained start values
  double coords_scale[MAXD];
  double obs_scale;
  int obs_mean_local;          // 1: subtract the per-expert mean of obs (obs_mean="local")
  double* obs_mean_out;        // [E] or nullptr
};

struct OptOut {
  double* theta;   // [E][MAXP]
  double* fobj;    // [E]  final -LML as seen by the optimiser
  int* status;     // [E]  LbStatus
  int* nit;        // [E]
  int* nfev;       // [E]
};

struct SlotAux {
  int* slot_expert;     // [S]
  LbfgsState* states;   // [S]
  int* queue_head;      // [1]
};

// CTA-wide: copy expert `e` into slot `s` (M1: coords /= coords_scale; obs = (obs - mean)/scale)
__device__ inline void load_expert_into_slot(const SlotCtx& c, int s, int e, const BatchIn& b, double* red) {
  const long long off = b.offsets[e];
  const int n = (int)(b.offsets[e + 1] - off);
  double mean = 0.0;
  if (b.obs_mean_local) {
    double v[1] = {0.0};
    for (int idx = threadIdx.x; idx < n; idx += NTHREADS) v[0] += b.obs[off + idx];
    block_sum<1>(v, red);
    if (threadIdx.x == 0) red[NTHREADS / 32] = v[0] / (double)n;
    __syncthreads();
    mean = red[NTHREADS / 32];
    __syncthreads();
  }
  double* ys = c.yobs + (long)s * c.npmax;
  for (int idx = threadIdx.x; idx < n; idx += NTHREADS) ys[idx] = (b.obs[off + idx] - mean) / b.obs_scale;
  double* cs = c.coords + (long)s * MAXD * c.npmax;
  for (int idx = threadIdx.x; idx < n * b.D; idx += NTHREADS) {
    const int row = idx / b.D, d = idx % b.D;
    cs[(long)d * c.npmax + row] = b.coords[(off + row) * b.D + d] / b.coords_scale[d];
  }
  if (threadIdx.x == 0) {
    c.n[s] = n;
    c.nb[s] = n / TB + 1;
    c.fail[s] = 0;
    if (b.obs_mean_out) b.obs_mean_out[e] = mean;
  }
}

// grid (S): initial fill.  Slot s takes expert order[first + s]; theta from theta_src[e].
// with_opt: initialise the optimiser state at u0 = inv(theta0) and set theta = fwd(u0).
__global__ void __launch_bounds__(NTHREADS) k_slot_init(SlotCtx c, SlotAux a, BatchIn b, TransformSpec tr,
                                                        int first, int count, int with_opt) {
  __shared__ double red[NTHREADS / 32 + 1];
  const int s = blockIdx.x;
  if (s >= count) {
    if (threadIdx.x == 0) c.active[s] = 0;
    return;
  }
  const int e = b.order ? b.order[first + s] : first + s;
  load_expert_into_slot(c, s, e, b, red);
  if (threadIdx.x == 0) {
    c.active[s] = 1;
    a.slot_expert[s] = e;
    double* th = c.theta + s * MAXP;
    for (int p = 0; p < tr.np; ++p) th[p] = b.theta0[(long)e * MAXP + p];
    if (with_opt) {
      double u0[MAXP];
      for (int k = 0; k < tr.nfree; ++k) {
        const int p = tr.free_idx[k];
        u0[k] = tr_inv(tr, p, th[p]);
        th[p] = tr_fwd(tr, p, u0[k]);
      }
      lb_init(a.states[s], u0, tr.nfree);
    }
  }
}

// grid (S): consume (f, g_theta) of the round, advance L-BFGS, emit results / refill the slot.
__global__ void __launch_bounds__(NTHREADS) k_opt_step(SlotCtx c, SlotAux a, BatchIn b, TransformSpec tr,
                                                       LbfgsOpts o, OptOut out) {
  __shared__ double red[NTHREADS / 32 + 1];
  __shared__ int sh_done, sh_next;
  const int s = blockIdx.x;
  if (!c.active[s]) return;
  if (threadIdx.x == 0) {
    LbfgsState& st = a.states[s];
    double* th = c.theta + s * MAXP;
    const double f = c.fout[s];
    double gu[MAXP];
    for (int k = 0; k < tr.nfree; ++k) {
      const int p = tr.free_idx[k];
      gu[k] = c.gout[s * MAXP + p] * tr_dfwd(tr, p, st.x[k]);
    }
    lbfgs_tell(st, o, f, gu);
    for (int k = 0; k < tr.nfree; ++k) {
      const int p = tr.free_idx[k];
      th[p] = tr_fwd(tr, p, st.x[k]);
    }
    sh_done = 0;
    if (st.status != LB_RUNNING) {
      const int e = a.slot_expert[s];
      for (int p = 0; p < tr.np; ++p) out.theta[(long)e * MAXP + p] = th[p];
      out.fobj[e] = st.f;
      out.status[e] = st.status;
      out.nit[e] = st.nit;
      out.nfev[e] = st.nfev;
      sh_done = 1;
      sh_next = atomicAdd(a.queue_head, 1);
    }
  }
  __syncthreads();
  if (!sh_done) return;
  const int nx = sh_next;
  if (nx >= b.E) {
    if (threadIdx.x == 0) c.active[s] = 0;
    return;
  }
  const int e = b.order ? b.order[nx] : nx;
  load_expert_into_slot(c, s, e, b, red);
  if (threadIdx.x == 0) {
    a.slot_expert[s] = e;
    double* th = c.theta + s * MAXP;
    for (int p = 0; p < tr.np; ++p) th[p] = b.theta0[(long)e * MAXP + p];
    double u0[MAXP];
    for (int k = 0; k < tr.nfree; ++k) {
      const int p = tr.free_idx[k];
      u0[k] = tr_inv(tr, p, th[p]);
      th[p] = tr_fwd(tr, p, u0[k]);
    }
    lb_init(a.states[s], u0, tr.nfree);
  }
}

// grid (S): copy per-slot evaluation results to per-expert outputs
__global__ void k_eval_scatter(SlotCtx c, SlotAux a, int count, double* fout, double* gout, int np) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
  const int e = a.slot_expert[s];
  if (fout) fout[e] = c.fout[s];
  if (gout)
    for (int p = 0; p < np; ++p) gout[(long)e * MAXP + p] = c.gout[s * MAXP + p];
}

// prediction inputs: grid (S) copy prediction coords of the slot's expert into [S][MAXD][ppmax]
__global__ void __launch_bounds__(NTHREADS) k_pred_load(SlotAux a, int count, int D, const double* pcoords,
                                                        const long long* poffsets, const double* coords_scale4,
                                                        double* pslot, int* np, int ppmax) {
  const int s = blockIdx.x;
  if (s >= count) return;
  const int e = a.slot_expert[s];
  const long long off = poffsets[e];
  const int P = (int)(poffsets[e + 1] - off);
  double* ps = pslot + (long)s * MAXD * ppmax;
  for (int idx = threadIdx.x; idx < P * D; idx += NTHREADS) {
    const int row = idx / D, d = idx % D;
    ps[(long)d * ppmax + row] = pcoords[(off + row) * D + d] / coords_scale4[d];
  }
  if (threadIdx.x == 0) np[s] = P;
}

// grid (S): scatter predictions back to the CSR outputs
__global__ void __launch_bounds__(NTHREADS) k_pred_scatter(SlotCtx c, SlotAux a, int count, const long long* poffsets,
                                                           const double* fmean_s, const double* fvar_s, int ppmax,
                                                           double* fmean, double* fvar, double* yvar, double* fobj) {
  const int s = blockIdx.x;
  if (s >= count) return;
  const int e = a.slot_expert[s];
  const long long off = poffsets[e];
  const int P = (int)(poffsets[e + 1] - off);
  const double nvar = c.theta[s * MAXP + c.D + 1];
  for (int idx = threadIdx.x; idx < P; idx += NTHREADS) {
    const double fv = fvar_s[(long)s * ppmax + idx];
    fmean[off + idx] = fmean_s[(long)s * ppmax + idx];
    fvar[off + idx] = fv;
    yvar[off + idx] = fv + nvar;
  }
  if (threadIdx.x == 0 && fobj) fobj[e] = c.fout[s];
}

}  // namespace gpsat
