// gpsat_b200: post-processing kernels either side of the hot path.
//   k_gauss_smooth    GPSat/postprocessing.py:22-52  gaussian_2d_weight (numba, target='parallel'):
//                     NaN-skipping Gaussian-weighted average of a hyper-parameter field, evaluated for every
//                     query point against all rows of its segment (smooth_hyperparameters,
//                     postprocessing.py:96-380, calls it once per (t, _dim_*) slice with x0 = x; here all
//                     slices go in ONE launch)
//   k_weighted_groups GPSat/utils.py:2081-2214      get_weighted_values, and the Gaussian glue of overlapping
//                     expert predictions (postprocessing.py:447-577): per-group sums of w and w*v with
//                     w = exp(-(|ref - to|^2 / l^2) / 2)
// Both are HBM/L2-bound streaming reductions: one CTA per output row, coalesced column reads, fp64 block sums.
#pragma once
#include "common.cuh"

namespace gpsat {

// grid (n_query), 256 threads.  seg_of_q[i] = segment of query i; rows of segment g are [seg_off[g], seg_off[g+1])
__global__ void __launch_bounds__(256) k_gauss_smooth(const double* __restrict__ qx, const double* __restrict__ qy,
                                                      const int* __restrict__ seg_of_q,
                                                      const double* __restrict__ x, const double* __restrict__ y,
                                                      const double* __restrict__ vals,
                                                      const long long* __restrict__ seg_off, double lx, double ly,
                                                      double vmin, double vmax, int clip_min, int clip_max,
                                                      double* __restrict__ out) {
  __shared__ double red[2 * 8];
  const long i = blockIdx.x;
  const int g = seg_of_q[i];
  const long long j0 = seg_off[g], j1 = seg_off[g + 1];
  const double x0 = qx[i], y0 = qy[i];
  double v[2] = {0.0, 0.0};   // w*val, w
  for (long long j = j0 + threadIdx.x; j < j1; j += 256) {
    double val = vals[j];
    if (val != val) continue;                     // skip NaN (postprocessing.py:41-44)
    if (clip_max && val > vmax) val = vmax;       // postprocessing.py:283-287 (applied before the weights)
    if (clip_min && val < vmin) val = vmin;
    const double dx = __ddiv_rn(x[j] - x0, lx), dy = __ddiv_rn(y[j] - y0, ly);
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double w = exp(-d2 / 2.0);
    v[0] += w * val;
    v[1] += w;
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) out[i] = (v[1] == 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : v[0] / v[1];
}

// grid (G), 256 threads: group g = rows order[off[g] .. off[g+1]) of the source frame.
// ref / to: [nd][n] column-major coordinates; vals [ncol][n]; out [ncol + 1][G] = sum(w v) / sum(w) per column,
// last row = sum(w)
__global__ void __launch_bounds__(256) k_weighted_groups(const double* __restrict__ ref, const double* __restrict__ to,
                                                         int nd, const double* __restrict__ vals, long n, int ncol,
                                                         const long long* __restrict__ order,
                                                         const long long* __restrict__ off, double l2, long G,
                                                         double* __restrict__ out) {
  __shared__ double red[2 * 8];
  const long g = blockIdx.x;
  const long long j0 = off[g], j1 = off[g + 1];
  for (int c0 = 0; c0 < ncol || c0 == 0; c0 += 1) {
    double v[2] = {0.0, 0.0};
    for (long long j = j0 + threadIdx.x; j < j1; j += 256) {
      const long long r = order[j];
      double d = 0.0;
      for (int k = 0; k < nd; ++k) {
        const double e = ref[(long)k * n + r] - to[(long)k * n + r];
        d = __dadd_rn(d, __dmul_rn(e, e));
      }
      const double w = exp(-__ddiv_rn(d, l2) / 2.0);
      if (ncol > 0) v[0] += w * vals[(long)c0 * n + r];
      v[1] += w;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) {
      if (ncol > 0) out[(long)c0 * G + g] = v[0] / v[1];
      if (c0 == 0) out[(long)ncol * G + g] = v[1];
    }
    __syncthreads();
    if (ncol == 0) break;
  }
}

}  // namespace gpsat
