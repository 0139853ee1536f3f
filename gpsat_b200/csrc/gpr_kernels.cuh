// gpsat_b200: batched exact-GPR kernels over "slots" (one resident expert per slot).
//
// Per objective evaluation for every active slot:
//   potrf  : left-looking blocked Cholesky of the AUGMENTED matrix [[K_y, y], [y', *]] with the
//            kernel matrix generated on the fly in the epilogue (K is never materialised);
//            row N of the factor is a' = (L^-1 y)'  -> quadratic form for free.
//   trtri  : X = L_aug^-1 by block forward substitution; row N of X is -alpha' = -(K_y^-1 y)'.
//   lauum  : tiles of X'X = K_y^-1 + alpha alpha' are formed in registers and contracted at once
//            with dK/dtheta generated on the fly:  G_k = sum_ij (K^-1 - aa')_ij dK_ij/dtheta_k
//            (K^-1 is never written to memory).
// Prediction (row F1): A = X K_xp accumulated tile by tile with a fused column sum of squares;
// the augmented row yields the posterior mean.
#pragma once
#include "gemm_core.cuh"

namespace gpsat {

constexpr int NG = MAXP;                         // gradient partials per tile
constexpr int SMEM_TILES = 4;                    // 2 stages x {A, B}
constexpr int AUX_DOUBLES = 2 * MAXD * TB + 2 * TB + 64;   // coords i/j, vectors, reduction scratch
constexpr int SMEM_BYTES = (SMEM_TILES * TILE_ELEMS + AUX_DOUBLES) * 8;

struct SlotCtx {
  int S, D, kid, nbmax, npmax, ntmax;
  long tile_stride;         // doubles per slot in Lt / Xt  (= ntmax * TILE_ELEMS)
  double* Lt;               // [S][ntmax][4096] packed lower tiles of L_aug
  double* Xt;               // [S][ntmax][4096] packed lower tiles of X = L_aug^-1
  double* coords;           // [S][MAXD][npmax]  coordinates / coords_scale
  double* yobs;             // [S][npmax]        (obs - mean) / scale
  int* n;                   // [S] observations per slot
  int* nb;                  // [S] 64-blocks of the augmented matrix = n/64 + 1
  int* active;              // [S]
  double* theta;            // [S][MAXP] lengthscales[D], kernel variance, likelihood variance
  double* logdet_part;      // [S][nbmax]
  double* gpart;            // [S][ntmax][NG]
  int* fail;                // [S] set when a pivot is not positive
  double* fout;             // [S]  -LML
  double* gout;             // [S][MAXP] d(-LML)/dtheta (constrained parameters)
};

__device__ __forceinline__ double* tile_ptr(double* base, int i, int j) {
  return base + tri_index(i, j) * TILE_ELEMS;
}

// stage the (scaled-by-1/l) coordinates of block `blk` into dst[MAXD][64]
__device__ __forceinline__ void stage_coords(double* dst, const double* coords_slot, int npmax, int D,
                                             const double* th, int blk, int N) {
  for (int t = threadIdx.x; t < D * TB; t += NTHREADS) {
    const int d = t / TB, m = t % TB, g = blk * TB + m;
    dst[d * TB + m] = (g < N) ? coords_slot[(long)d * npmax + g] / th[d] : 0.0;
  }
}

// ------------------------------------------------------------------------------------
// 64x64 diagonal block: Cholesky + triangular inverse in shared memory
// a: [64][65] (lower part valid), inv: [64][68], dg: [64].  All NTHREADS threads call.
// Global indices >= N (augmented row and padding) get a forced unit pivot.
// ------------------------------------------------------------------------------------
constexpr int LDA = 65;
constexpr int LDI = 68;

__device__ __forceinline__ void potf2_trtri_64(double* a, double* inv, double* dg, int g0, int N, int* fail_flag) {
  const int tid = threadIdx.x;
  const int rr = tid & 63, cg = tid >> 6;
  for (int c = 0; c < TB; ++c) {
    __syncthreads();
    double d = a[c * LDA + c];
    if (g0 + c >= N) d = 1.0;
    if (!(d > 0.0)) {
      if (tid == 0) *fail_flag = 1;
      d = 1.0;
    }
    const double piv = sqrt(d);
    if (tid < TB) {
      if (tid > c) a[tid * LDA + c] *= (1.0 / piv);
      else if (tid == c) dg[c] = piv;
    }
    __syncthreads();
    const double lrc = a[rr * LDA + c];
    for (int cc = c + 1 + cg; cc <= rr; cc += 4) a[rr * LDA + cc] -= lrc * a[cc * LDA + c];
  }
  __syncthreads();
  // inverse: 4 threads per column, uniform row loop
  const int col = tid >> 2, part = tid & 3;
  for (int r = 0; r < TB; ++r) {
    double sum = 0.0;
    for (int k = part; k < r; k += 4) sum += a[r * LDA + k] * inv[k * LDI + col];
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (part == 0) inv[r * LDI + col] = (r >= col) ? (((r == col) ? 1.0 : 0.0) - sum) / dg[r] : 0.0;
    __syncwarp();
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------
// potrf step j, part 1: C_ij = K_aug(i, j) - sum_{k<j} L_ik L_jk'   for all i >= j.
// The diagonal CTA (i == j) also factorises C_jj -> L_jj, L_jj^-1 (-> X_jj) and log-det part.
// grid (nbmax - j, S)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) k_potrf_update(SlotCtx c, int j) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  if (!c.active[s]) return;
  const int nb = c.nb[s];
  const int i = j + blockIdx.x;
  if (i >= nb) return;
  const int N = c.n[s];
  double* Lt = c.Lt + (long)s * c.tile_stride;
  FragCoord fc;
  Acc acc;
  acc.zero();
  gemm_pipeline<false, false>(
      acc, smem, 0, j, [&](int k) { return tile_ptr(Lt, i, k); }, [&](int k) { return tile_ptr(Lt, j, k); }, fc);

  double* aux = smem + SMEM_TILES * TILE_ELEMS;
  double* xi = aux;
  double* xj = aux + MAXD * TB;
  double* yj = aux + 2 * MAXD * TB;
  const double* th = c.theta + s * MAXP;
  const double* cs = c.coords + (long)s * MAXD * c.npmax;
  stage_coords(xi, cs, c.npmax, c.D, th, i, N);
  stage_coords(xj, cs, c.npmax, c.D, th, j, N);
  if (threadIdx.x < TB) {
    const int g = j * TB + threadIdx.x;
    yj[threadIdx.x] = (g < N) ? c.yobs[(long)s * c.npmax + g] : 0.0;
  }
  __syncthreads();
  const double kvar = th[c.D], nvar = th[c.D + 1];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int m = fc.row(mi), gi = i * TB + m;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = fc.col(ni) + e, gj = j * TB + n;
        double val;
        if (gi < N && gj < N) {
          double r2 = 0.0;
          for (int d = 0; d < c.D; ++d) {
            const double df = xi[d * TB + m] - xj[d * TB + n];
            r2 += df * df;
          }
          val = kern_value(c.kid, r2, kvar);
          if (gi == gj) val += nvar;
        } else if (gi == N && gj < N) {
          val = yj[n];
        } else {
          val = (gi == gj) ? 1.0 : 0.0;
        }
        acc.c[mi][ni][e] = val - acc.c[mi][ni][e];
      }
    }
  }
  if (i != j) {
    store_acc_swizzled(tile_ptr(Lt, i, j), acc, fc);
    return;
  }
  // ---- diagonal block ----
  double* a = smem;                      // [64][65]
  double* inv = smem + 2 * TILE_ELEMS;   // [64][68]
  double* dg = yj;                       // reuse
  __syncthreads();
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) a[fc.row(mi) * LDA + fc.col(ni) + e] = acc.c[mi][ni][e];
  potf2_trtri_64(a, inv, dg, j * TB, N, c.fail + s);
  double* Ld = tile_ptr(Lt, j, j);
  double* Xd = tile_ptr(c.Xt + (long)s * c.tile_stride, j, j);
  for (int t = threadIdx.x; t < TILE_ELEMS; t += NTHREADS) {
    const int r = t >> 6, cc = t & 63;
    const double lv = (cc < r) ? a[r * LDA + cc] : ((cc == r) ? dg[r] : 0.0);
    Ld[swz(r, cc)] = lv;
    Xd[swz(r, cc)] = inv[r * LDI + cc];
  }
  if (threadIdx.x < 32) {
    double ld = 0.0;
    for (int k = threadIdx.x; k < TB; k += 32)
      if (j * TB + k < N) ld += log(dg[k]);
    ld = warp_sum(ld);
    if (threadIdx.x == 0) c.logdet_part[s * c.nbmax + j] = ld;
  }
}

// potrf step j, part 2: L_ij = C_ij * L_jj^-T  for i > j.   grid (nbmax - j - 1, S)
__global__ void __launch_bounds__(NTHREADS, 1) k_potrf_trsm(SlotCtx c, int j) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  if (!c.active[s]) return;
  const int i = j + 1 + blockIdx.x;
  if (i >= c.nb[s]) return;
  double* Lt = c.Lt + (long)s * c.tile_stride;
  double* Xt = c.Xt + (long)s * c.tile_stride;
  FragCoord fc;
  Acc acc;
  acc.zero();
  gemm_pipeline<false, false>(
      acc, smem, 0, 1, [&](int) { return tile_ptr(Lt, i, j); }, [&](int) { return tile_ptr(Xt, j, j); }, fc);
  store_acc_swizzled(tile_ptr(Lt, i, j), acc, fc);
}

// trtri step sd >= 1: X_{j+sd, j} = -L_ii^-1 * sum_{k=j}^{i-1} L_ik X_kj.   grid (nbmax - sd, S)
__global__ void __launch_bounds__(NTHREADS, 1) k_trtri_step(SlotCtx c, int sd) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  if (!c.active[s]) return;
  const int j = blockIdx.x, i = j + sd;
  if (i >= c.nb[s]) return;
  double* Lt = c.Lt + (long)s * c.tile_stride;
  double* Xt = c.Xt + (long)s * c.tile_stride;
  FragCoord fc;
  Acc acc;
  acc.zero();
  gemm_pipeline<false, true>(
      acc, smem, j, i, [&](int k) { return tile_ptr(Lt, i, k); }, [&](int k) { return tile_ptr(Xt, k, j); }, fc);
  // T -> shared (B operand, [kk][n]); L_ii^-1 -> shared (A operand)
  double* Ts = smem;
  double* Ls = smem + TILE_ELEMS;
  load_tile_async(Ls, tile_ptr(Xt, i, i));
  cp_async_commit();
  store_acc_swizzled(Ts, acc, fc);
  cp_async_wait<0>();
  __syncthreads();
  Acc acc2;
  acc2.zero();
  mma_tile<false, true>(acc2, Ls, Ts, fc);
  store_acc_swizzled(tile_ptr(Xt, i, j), acc2, fc, -1.0);
}

// lauum + trace: tile (i, j), i >= j of X'X and the gradient contraction.   grid (ntmax, S)
__global__ void __launch_bounds__(NTHREADS, 1) k_lauum_trace(SlotCtx c) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  if (!c.active[s]) return;
  const int nb = c.nb[s];
  const int t = blockIdx.x;
  int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((long)(i + 1) * (i + 2) / 2 <= t) ++i;
  while ((long)i * (i + 1) / 2 > t) --i;
  const int j = t - (int)((long)i * (i + 1) / 2);
  if (i >= nb) return;
  const int N = c.n[s];
  double* Xt = c.Xt + (long)s * c.tile_stride;
  FragCoord fc;
  Acc acc;
  acc.zero();
  gemm_pipeline<true, true>(
      acc, smem, i, nb, [&](int k) { return tile_ptr(Xt, k, i); }, [&](int k) { return tile_ptr(Xt, k, j); }, fc);

  double* aux = smem + SMEM_TILES * TILE_ELEMS;
  double* xi = aux;
  double* xj = aux + MAXD * TB;
  double* ai = aux + 2 * MAXD * TB;
  double* aj = ai + TB;
  double* red = aj + TB;
  const double* th = c.theta + s * MAXP;
  const double* cs = c.coords + (long)s * MAXD * c.npmax;
  stage_coords(xi, cs, c.npmax, c.D, th, i, N);
  stage_coords(xj, cs, c.npmax, c.D, th, j, N);
  const int bN = nb - 1, rN = N - bN * TB;
  if (threadIdx.x < TB) {
    ai[threadIdx.x] = -tile_ptr(Xt, bN, i)[swz(rN, threadIdx.x)];
  } else if (threadIdx.x < 2 * TB) {
    const int m = threadIdx.x - TB;
    aj[m] = -tile_ptr(Xt, bN, j)[swz(rN, m)];
  }
  __syncthreads();
  const double kvar = th[c.D];
  double g[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) g[k] = 0.0;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int m = fc.row(mi), gi = i * TB + m;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = fc.col(ni) + e, gj = j * TB + n;
        if (gi < N && gj < N) {
          const double W = acc.c[mi][ni][e] - 2.0 * ai[m] * aj[n];
          double r2 = 0.0, d2[MAXD];
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            d2[d] = 0.0;
            if (d < c.D) {
              const double df = xi[d * TB + m] - xj[d * TB + n];
              d2[d] = df * df;
              r2 += d2[d];
            }
          }
          double kv, hv;
          kern_eval(c.kid, r2, kvar, kv, hv);
          const double wh = W * hv;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) g[d] += wh * d2[d];
          g[MAXD] += W * kv;
          if (gi == gj) g[MAXD + 1] += W;
        }
      }
    }
  }
  const double wgt = (i != j) ? 2.0 : 1.0;
  block_sum<NG>(g, red);
  if (threadIdx.x == 0) {
    double* gp = c.gpart + ((long)s * c.ntmax + t) * NG;
#pragma unroll
    for (int k = 0; k < NG; ++k) gp[k] = wgt * g[k];
  }
}

// finalize: -LML and d(-LML)/dtheta per slot.  grid (S), NTHREADS threads
__global__ void __launch_bounds__(NTHREADS) k_finalize(SlotCtx c, int with_grad) {
  __shared__ double red[(NG + 2) * (NTHREADS / 32)];
  const int s = blockIdx.x;
  if (!c.active[s]) return;
  const int N = c.n[s], nb = c.nb[s];
  const int bN = nb - 1, rN = N - bN * TB;
  const double* Lt = c.Lt + (long)s * c.tile_stride;
  double v[NG + 2];
#pragma unroll
  for (int k = 0; k < NG + 2; ++k) v[k] = 0.0;
  for (int idx = threadIdx.x; idx < N; idx += NTHREADS) {
    const int k = idx >> 6, cc = idx & 63;
    const double a = Lt[tri_index(bN, k) * TILE_ELEMS + swz(rN, cc)];
    v[NG] += a * a;
  }
  for (int k = threadIdx.x; k < nb; k += NTHREADS) v[NG + 1] += c.logdet_part[s * c.nbmax + k];
  if (with_grad) {
    const int nt = nb * (nb + 1) / 2;
    for (int t = threadIdx.x; t < nt; t += NTHREADS) {
      const double* gp = c.gpart + ((long)s * c.ntmax + t) * NG;
#pragma unroll
      for (int k = 0; k < NG; ++k) v[k] += gp[k];
    }
  }
  block_sum<NG + 2>(v, red);
  if (threadIdx.x == 0) {
    const double* th = c.theta + s * MAXP;
    double f = 0.5 * v[NG] + v[NG + 1] + 0.5 * N * 1.8378770664093453;
    if (c.fail[s]) f = INFINITY;
    c.fail[s] = 0;  // consumed: ready for the next evaluation
    c.fout[s] = f;
    double* go = c.gout + s * MAXP;
    for (int d = 0; d < c.D; ++d) go[d] = 0.5 * v[d] / th[d];
    go[c.D] = 0.5 * v[MAXD] / th[c.D];
    go[c.D + 1] = 0.5 * v[MAXD + 1];
  }
}

// ------------------------------------------------------------------------------------
// Prediction.  Persistent CTAs walk work items (slot, 64-wide block of prediction points):
//   phase 1: K_xp tiles for the item -> CTA-private global scratch (L2 resident)
//   phase 2: A_i = sum_{k<=i} X_ik K_xp,k ; column sums of A^2 over rows < N; row N -> -mean
// ------------------------------------------------------------------------------------
struct PredCtx {
  int ppmax;                 // padded prediction points per slot
  const double* pcoords;     // [S][MAXD][ppmax] prediction coords / coords_scale
  const int* np;             // [S]
  const int* item_slot;      // [n_items]
  const int* item_pb;        // [n_items]
  int n_items;
  double* scratch;           // [gridDim.x][nbmax][4096]
  double* fmean;             // [S][ppmax]
  double* fvar;              // [S][ppmax]
};

__global__ void __launch_bounds__(NTHREADS, 1) k_predict(SlotCtx c, PredCtx p) {
  extern __shared__ __align__(128) double smem[];
  double* aux = smem + SMEM_TILES * TILE_ELEMS;
  double* xo = aux;                  // obs coords block  [MAXD][64]
  double* xp = aux + MAXD * TB;      // pred coords block [MAXD][64]
  double* colsq = aux + 2 * MAXD * TB;   // [64]
  double* meanv = colsq + TB;            // [64]
  double* red = meanv + TB;              // [4][64] needs 256 doubles -> lives in tile area when used
  (void)red;
  FragCoord fc;
  double* scr = p.scratch + (long)blockIdx.x * c.nbmax * TILE_ELEMS;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int s = p.item_slot[item], pb = p.item_pb[item];
    const int N = c.n[s], nb = c.nb[s], P = p.np[s];
    const double* th = c.theta + s * MAXP;
    const double* cs = c.coords + (long)s * MAXD * c.npmax;
    const double* ps = p.pcoords + (long)s * MAXD * p.ppmax;
    double* Xt = c.Xt + (long)s * c.tile_stride;
    const double kvar = th[c.D];
    __syncthreads();
    stage_coords(xp, ps, p.ppmax, c.D, th, pb, P);
    // ---- phase 1 ----
    for (int k = 0; k < nb; ++k) {
      __syncthreads();
      stage_coords(xo, cs, c.npmax, c.D, th, k, N);
      __syncthreads();
      double* T = scr + (long)k * TILE_ELEMS;
      for (int t = threadIdx.x; t < TILE_ELEMS; t += NTHREADS) {
        const int kk = t >> 6, n = t & 63;
        double val = 0.0;
        if (k * TB + kk < N && pb * TB + n < P) {
          double r2 = 0.0;
          for (int d = 0; d < c.D; ++d) {
            const double df = xo[d * TB + kk] - xp[d * TB + n];
            r2 += df * df;
          }
          val = kern_value(c.kid, r2, kvar);
        }
        T[swz(kk, n)] = val;
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    double csq[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) csq[ni][0] = csq[ni][1] = 0.0;
    const int bN = nb - 1, rN = N - bN * TB;
    for (int i = 0; i < nb; ++i) {
      Acc acc;
      acc.zero();
      gemm_pipeline<false, true>(
          acc, smem, 0, i + 1, [&](int k) { return tile_ptr(Xt, i, k); },
          [&](int k) { return scr + (long)k * TILE_ELEMS; }, fc);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int gi = i * TB + fc.row(mi);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if (gi < N) {
            csq[ni][0] += acc.c[mi][ni][0] * acc.c[mi][ni][0];
            csq[ni][1] += acc.c[mi][ni][1] * acc.c[mi][ni][1];
          } else if (i == bN && fc.row(mi) == rN) {
            meanv[fc.col(ni)] = -acc.c[mi][ni][0];
            meanv[fc.col(ni) + 1] = -acc.c[mi][ni][1];
          }
        }
      }
    }
    // reduce column sums: over q (lanes sharing r) then over the 4 M-warps
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = csq[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        csq[ni][e] = v;
      }
    double* part = smem;  // [4][64] (tile area is idle here: pipeline ended with a barrier)
    if (fc.q == 0) {
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        part[fc.wm * TB + fc.col(ni)] = csq[ni][0];
        part[fc.wm * TB + fc.col(ni) + 1] = csq[ni][1];
      }
    }
    __syncthreads();
    if (threadIdx.x < TB) {
      const int n = threadIdx.x, gp = pb * TB + n;
      if (gp < P) {
        const double ssq = (part[n] + part[TB + n]) + (part[2 * TB + n] + part[3 * TB + n]);
        p.fvar[(long)s * p.ppmax + gp] = kvar - ssq;
        p.fmean[(long)s * p.ppmax + gp] = meanv[n];
      }
    }
    (void)colsq;
  }
}

// dense row-major kernel matrix for parity tests / HBM roofline of the kernel build (row K1)
// K[i][j] = k(x_i, x2_j) (+ nvar on the diagonal when add_noise).  grid (ceil(n2/64), ceil(n1/16))
__global__ void k_kernel_matrix(const double* __restrict__ X1, int n1, const double* __restrict__ X2, int n2,
                                int D, int kid, const double* __restrict__ theta, int add_noise,
                                double* __restrict__ K) {
  __shared__ double x1s[16][MAXD], x2s[64][MAXD];
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 64;
  const int tid = threadIdx.y * 64 + threadIdx.x;
  for (int t = tid; t < 16 * D; t += 256) {
    const int r = t / D, d = t % D;
    x1s[r][d] = (i0 + r < n1) ? X1[(long)(i0 + r) * D + d] / theta[d] : 0.0;
  }
  for (int t = tid; t < 64 * D; t += 256) {
    const int r = t / D, d = t % D;
    x2s[r][d] = (j0 + r < n2) ? X2[(long)(j0 + r) * D + d] / theta[d] : 0.0;
  }
  __syncthreads();
  const int j = j0 + threadIdx.x;
  if (j >= n2) return;
  for (int rr = threadIdx.y; rr < 16; rr += 4) {
    const int i = i0 + rr;
    if (i >= n1) break;
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = x1s[rr][d] - x2s[threadIdx.x][d];
      r2 += df * df;
    }
    double v = kern_value(kid, r2, theta[D]);
    if (add_noise && i == j) v += theta[D + 1];
    K[(long)i * n2 + j] = v;
  }
}

// unpack packed swizzled lower tiles of one slot into a dense row-major (npad x npad) matrix
__global__ void k_unpack_tiles(const double* __restrict__ tiles, int nb, double* __restrict__ dense) {
  const int i = blockIdx.y, j = blockIdx.x;
  const int npad = nb * TB;
  for (int t = threadIdx.x; t < TILE_ELEMS; t += blockDim.x) {
    const int r = t >> 6, cc = t & 63;
    double v = 0.0;
    if (j <= i) v = tiles[tri_index(i, j) * TILE_ELEMS + swz(r, cc)];
    dense[(long)(i * TB + r) * npad + j * TB + cc] = v;
  }
}

}  // namespace gpsat
