// gpsat_b200: observation / prediction-location selection.
//
// Bit-exact restatement of
//   DataLoader.local_data_select            (GPSat/dataloader.py:2352-2447)
//   PredictionLocations._max_dist_bool      (GPSat/prediction_locations.py:18-43)
// as an order-preserving two-pass stream compaction: pass 1 counts matches per expert, pass 2
// writes the matching row indices in ascending (source) order.  All distance arithmetic uses
// __dmul_rn / __dadd_rn so no FMA contraction can change a tie on the radius.
//   type 0: scalar window   x[col] <comp> (ref[col] + val)
//   type 1: KD-tree ball    sum_j (x_j - ref_j)^2 <= r*r            (inclusive)
//   type 2: max_dist        every (d_j*d_j) < r*r and sum_j d_j*d_j < r*r   (strict)
#pragma once
#include "common.cuh"

namespace gpsat {

constexpr int SEL_MAXTERMS = 8;
constexpr int SEL_MAXCOL = 4;

struct SelTerm {
  int type, ncol, comp, pad_;
  int col[SEL_MAXCOL];   // column index into the observation table
  int rcol[SEL_MAXCOL];  // column index into the expert (reference) table
  double val;
};
struct SelSpec {
  int nterms, pad_;
  SelTerm t[SEL_MAXTERMS];
};

__device__ __forceinline__ bool sel_cmp(int comp, double a, double b) {
  switch (comp) {
    case 0: return a >= b;
    case 1: return a > b;
    case 2: return a == b;
    case 3: return a < b;
    default: return a <= b;
  }
}

// obs: [ncols][n] column-major table; ref: this expert's row [nrefcols]
__device__ __forceinline__ bool sel_match(const SelSpec& sp, const double* __restrict__ obs, long n, long row,
                                          const double* ref) {
  bool ok = true;
  for (int k = 0; k < sp.nterms && ok; ++k) {
    const SelTerm& t = sp.t[k];
    if (t.type == 0) {
      ok = sel_cmp(t.comp, obs[(long)t.col[0] * n + row], __dadd_rn(ref[t.rcol[0]], t.val));
    } else {
      const double r2 = __dmul_rn(t.val, t.val);
      double d2 = 0.0;
      for (int j = 0; j < t.ncol; ++j) {
        const double d = __dadd_rn(obs[(long)t.col[j] * n + row], -ref[t.rcol[j]]);
        const double dd = __dmul_rn(d, d);
        if (t.type == 2 && !(dd < r2)) ok = false;
        d2 = __dadd_rn(d2, dd);
      }
      ok = ok && ((t.type == 1) ? (d2 <= r2) : (d2 < r2));
    }
  }
  return ok;
}

// grid (E), 256 threads.  fill == 0: counts[e] = matches.  fill == 1: idx[offsets[e] ...] = rows.
__global__ void __launch_bounds__(256) k_select(SelSpec sp, const double* __restrict__ obs, long n,
                                                const double* __restrict__ refs, int nrefcols, int fill,
                                                long long* counts, const long long* offsets, int* idx) {
  __shared__ double ref[16];
  __shared__ int wsum[8];
  __shared__ long long base_sh;
  const int e = blockIdx.x;
  if (threadIdx.x < nrefcols) ref[threadIdx.x] = refs[(long)e * nrefcols + threadIdx.x];
  if (threadIdx.x == 0) base_sh = fill ? offsets[e] : 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long total = 0;
  for (long r0 = 0; r0 < n; r0 += 256) {
    const long row = r0 + threadIdx.x;
    const bool m = (row < n) && sel_match(sp, obs, n, row, ref);
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (!fill) {
      if (lane == 0) total += __popc(bal);
      continue;
    }
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int v = wsum[w];
      if (w < warp) before += v;
      all += v;
    }
    if (m) idx[base_sh + before + __popc(bal & ((1u << lane) - 1u))] = (int)row;
    __syncthreads();
    if (threadIdx.x == 0) base_sh += all;
  }
  if (!fill) {
    // lane-0 partials -> block total
    if (lane == 0) wsum[warp] = (int)total;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < 8; ++w) t += wsum[w];
      counts[e] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Grid-bucketed variant: rows are binned once into square cells of edge >= the query radius on the
// two columns of the ball / max_dist term; an expert then tests only the rows of its 3 x 3 cell
// neighbourhood with the SAME predicate arithmetic (bit-exact), collects the matches in shared memory
// and sorts them, so the output is again in ascending source-row order.
// ------------------------------------------------------------------------------------------------
struct CellGrid {
  double x0, y0, inv_cell;
  int ncx, ncy, colx, coly, rcolx, rcoly;
};
__device__ __forceinline__ int cell_coord(double v, double v0, double inv_cell) {
  const double f = floor((v - v0) * inv_cell);
  return (int)fmax(-2.0, fmin(f, 2.0e9));
}
// counts[c] += 1 (hist) / order[start[c] + cursor[c]++] = row (scatter).  grid-stride over rows
__global__ void __launch_bounds__(256) k_cell_hist(CellGrid g, const double* __restrict__ tab, long n, int* counts) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cx = min(max(cell_coord(tab[(long)g.colx * n + i], g.x0, g.inv_cell), 0), g.ncx - 1);
    const int cy = min(max(cell_coord(tab[(long)g.coly * n + i], g.y0, g.inv_cell), 0), g.ncy - 1);
    atomicAdd(counts + (long)cy * g.ncx + cx, 1);
  }
}
__global__ void __launch_bounds__(256) k_cell_scatter(CellGrid g, const double* __restrict__ tab, long n,
                                                      const long long* __restrict__ start, int* cursor, int* order) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cx = min(max(cell_coord(tab[(long)g.colx * n + i], g.x0, g.inv_cell), 0), g.ncx - 1);
    const int cy = min(max(cell_coord(tab[(long)g.coly * n + i], g.y0, g.inv_cell), 0), g.ncy - 1);
    const long c = (long)cy * g.ncx + cx;
    order[start[c] + atomicAdd(cursor + c, 1)] = (int)i;
  }
}

// grid (E), 256 threads, dynamic smem: cap ints (fill mode; cap = power of two >= max count)
__global__ void __launch_bounds__(256) k_select_bucket(SelSpec sp, CellGrid g, const double* __restrict__ obs, long n,
                                                       const double* __restrict__ refs, int nrefcols,
                                                       const long long* __restrict__ start,
                                                       const int* __restrict__ order, int fill, int cap,
                                                       long long* counts, const long long* offsets, int* idx) {
  extern __shared__ int list[];
  __shared__ double ref[16];
  __shared__ int nfound;
  const int e = blockIdx.x;
  if (threadIdx.x < nrefcols) ref[threadIdx.x] = refs[(long)e * nrefcols + threadIdx.x];
  if (threadIdx.x == 0) nfound = 0;
  __syncthreads();
  const int ecx = cell_coord(ref[g.rcolx], g.x0, g.inv_cell), ecy = cell_coord(ref[g.rcoly], g.y0, g.inv_cell);
  int local = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int cy = ecy + dy;
    if (cy < 0 || cy >= g.ncy) continue;
    // rows clamped into the border cells may lie outside the nominal grid: widen the x-range there
    const int cx0 = max(ecx - 1, 0), cx1 = min(ecx + 1, g.ncx - 1);
    if (cx0 > cx1) continue;
    const long long k0 = start[(long)cy * g.ncx + cx0], k1 = start[(long)cy * g.ncx + cx1 + 1];
    for (long long k = k0 + threadIdx.x; k < k1; k += 256) {
      const int row = order[k];
      if (sel_match(sp, obs, n, row, ref)) {
        if (fill) list[atomicAdd(&nfound, 1)] = row;
        else ++local;
      }
    }
  }
  if (!fill) {
    atomicAdd(&nfound, local);
    __syncthreads();
    if (threadIdx.x == 0) counts[e] = nfound;
    return;
  }
  __syncthreads();
  const int m = nfound;
  for (int t = m + threadIdx.x; t < cap; t += 256) list[t] = 0x7fffffff;
  __syncthreads();
  for (int k2 = 2; k2 <= cap; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < cap; t += 256) {
        const int p = t ^ j;
        if (p > t) {
          const int a = list[t], b = list[p];
          const bool up = ((t & k2) == 0);
          if ((a > b) == up) { list[t] = b; list[p] = a; }
        }
      }
      __syncthreads();
    }
  const long long base = offsets[e];
  for (int t = threadIdx.x; t < m; t += 256) idx[base + t] = list[t];
}

// gather selected rows of a column-major table into the CSR batch layout the GPR entry points
// take: coords[total][D] row-major (from columns ccols[0..D-1]) and obs[total] (column ocol).
struct GatherCols {
  int D, ocol;
  int ccols[MAXD];
};
__global__ void __launch_bounds__(256) k_gather_rows(GatherCols gc, const double* __restrict__ table, long n,
                                                     const int* __restrict__ idx, long total,
                                                     double* __restrict__ coords, double* __restrict__ obs) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = idx[t];
    for (int d = 0; d < gc.D; ++d) coords[t * gc.D + d] = table[(long)gc.ccols[d] * n + row];
    if (obs) obs[t] = table[(long)gc.ocol * n + row];
  }
}

// prediction coordinates of every expert: selected rows of the prediction-location table for the
// columns it has (tcol[d] >= 0), the expert's own value for the others (tcol[d] < 0 -> refs[e][rcol[d]])
// (PredictionLocations._from_dataframe, prediction_locations.py:258-271).  grid (E)
struct PredCols {
  int D;
  int tcol[MAXD];
  int rcol[MAXD];
};
__global__ void __launch_bounds__(256) k_gather_pred(PredCols pc, const double* __restrict__ table, long n,
                                                     const double* __restrict__ refs, int nrefcols,
                                                     const long long* __restrict__ offsets,
                                                     const int* __restrict__ idx, double* __restrict__ out) {
  const int e = blockIdx.x;
  const long long o0 = offsets[e], o1 = offsets[e + 1];
  for (long long t = o0 + threadIdx.x; t < o1; t += blockDim.x) {
    const long row = idx[t];
    for (int d = 0; d < pc.D; ++d)
      out[t * pc.D + d] = (pc.tcol[d] >= 0) ? table[(long)pc.tcol[d] * n + row] : refs[(long)e * nrefcols + pc.rcol[d]];
  }
}

}  // namespace gpsat
