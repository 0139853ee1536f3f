// gpsat_b200: upstream binning kernel.
//   k_bin_accumulate   GPSat/dataprepper.py:230-407 DataPrep.bin_data -> scipy.stats.binned_statistic_2d(statistic =
//                      "mean" | "sum" | "count") for every by_cols group of DataPrep.bin_data_by (dataprepper.py:23-228)
//                      in ONE launch: bin numbers exactly as scipy's binned_statistic_dd assigns them
//                      (searchsorted(edges, v, side="right"), values that round onto the last edge go into the last
//                      bin), then fp64 atomic sums and integer counts per (group, x bin, y bin).
//   k_bin_spread       second pass for statistic = "std" | "min" | "max": with the per-bin sums and counts of the first
//                      pass known, accumulates sum (v - mean_bin)^2 (np.std's two-pass form, which is what scipy applies
//                      per bin) and the per-bin extrema (compare-and-swap on the fp64 bit pattern).
// HBM-bound scatter-reduce: 24-28 bytes read per observation, atomics resolved in L2.
#pragma once
#include "common.cuh"

namespace gpsat {

struct BinAxis {
  const double* edges;   // [n_edges] ascending
  int n_edges;
  double round_scale;    // 10^|decimal| of scipy's on-edge test
  int round_div;         // 1: around(v, decimal) = rint(v / scale) * scale (decimal < 0); 0: rint(v * scale) / scale
};

// scipy _bin_numbers: Ncount = digitize(v, edges); if v >= edges[-1] and around(v, decimal) == around(edges[-1], decimal):
// Ncount -= 1
__device__ __forceinline__ int bin_number(const BinAxis& ax, double v) {
  int lo = 0, hi = ax.n_edges;                  // first index with edges[idx] > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ax.edges[mid] > v) hi = mid; else lo = mid + 1;
  }
  const double e = ax.edges[ax.n_edges - 1];
  if (v >= e) {
    const double rv = ax.round_div ? rint(__ddiv_rn(v, ax.round_scale)) : rint(__dmul_rn(v, ax.round_scale));
    const double re = ax.round_div ? rint(__ddiv_rn(e, ax.round_scale)) : rint(__dmul_rn(e, ax.round_scale));
    if (rv == re) lo -= 1;
  }
  return lo;                                    // 0: below range, n_edges: above range
}

// grid-stride over rows.  sum / cnt: [n_groups][nx][ny] with nx = x.n_edges - 1, ny = y.n_edges - 1 (ny = 1 for 1-D)
__global__ void __launch_bounds__(256) k_bin_accumulate(const double* __restrict__ x, const double* __restrict__ y,
                                                        const double* __restrict__ vals,
                                                        const int* __restrict__ group, long long n, BinAxis ax,
                                                        BinAxis ay, int two_d, double* __restrict__ sum,
                                                        unsigned long long* __restrict__ cnt) {
  const int nx = ax.n_edges - 1, ny = two_d ? ay.n_edges - 1 : 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int bx = bin_number(ax, x[i]);
    if (bx < 1 || bx > nx) continue;
    int by = 1;
    if (two_d) {
      by = bin_number(ay, y[i]);
      if (by < 1 || by > ny) continue;
    }
    const long long g = group ? group[i] : 0;
    const long long b = (g * nx + (bx - 1)) * ny + (by - 1);
    atomicAdd(sum + b, vals[i]);
    atomicAdd(cnt + b, 1ULL);
  }
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(a);
  while (v < __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(a);
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}

// same row -> bin map as k_bin_accumulate.  sum / cnt: the first pass.  ssd, vmin, vmax: [n_groups][nx][ny] or nullptr
// (ssd zeroed, vmin = +inf, vmax = -inf by the caller).
__global__ void __launch_bounds__(256) k_bin_spread(const double* __restrict__ x, const double* __restrict__ y,
                                                    const double* __restrict__ vals, const int* __restrict__ group,
                                                    long long n, BinAxis ax, BinAxis ay, int two_d,
                                                    const double* __restrict__ sum,
                                                    const unsigned long long* __restrict__ cnt,
                                                    double* __restrict__ ssd, double* __restrict__ vmin,
                                                    double* __restrict__ vmax) {
  const int nx = ax.n_edges - 1, ny = two_d ? ay.n_edges - 1 : 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int bx = bin_number(ax, x[i]);
    if (bx < 1 || bx > nx) continue;
    int by = 1;
    if (two_d) {
      by = bin_number(ay, y[i]);
      if (by < 1 || by > ny) continue;
    }
    const long long g = group ? group[i] : 0;
    const long long b = (g * nx + (bx - 1)) * ny + (by - 1);
    const double v = vals[i];
    if (ssd) {
      const double d = v - __ddiv_rn(sum[b], (double)cnt[b]);
      atomicAdd(ssd + b, __dmul_rn(d, d));
    }
    if (vmin) atomic_min_f64(vmin + b, v);
    if (vmax) atomic_max_f64(vmax + b, v);
  }
}

}  // namespace gpsat
