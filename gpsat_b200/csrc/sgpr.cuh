// gpsat_b200: batched sparse GPR (Titsias' collapsed bound), SURVEY 8a row SG1
// (GPSat/models/gpflow_models.py:666-901 -> gpflow.models.SGPR.elbo / predict_f).
//
// Per expert: data X (N x D), y (N), inducing points Z (M x D), theta = (l, s_f^2, s_n^2), beta = 1/s_n^2.
//   Kuu = K(Z,Z) + 1e-6 I = L L'      (Z-slot pool `cz`: the exact-GPR kernels factorise and invert it)
//   A'  = L^-1 Kuf                    (M x N)           B = I + beta A' A'^T = LB LB'   (slot pool `cb`)
//   c   = LB^-1 (beta A' y)           (augmented row of LB)
//   ELBO = -N/2 log 2pi - sum log LB_ii - N/2 log s_n^2 - beta/2 y'y + c'c/2 - beta/2 N s_f^2 + beta/2 tr(A'A'^T)
// Gradient (analytic, whitened):  u = B^-1 A'y,  w = L^-T u,  s = A'^T u,  r = y - beta s,
//   dF/dKuf = L^-T [beta (I - B^-1)] A' + beta^2 w r^T            =: T A' + beta^2 w r^T
//   dF/dKuu = 1/2 L^-T (2I - B^-1 - B) L^-1 - 1/2 beta^2 w w^T
// contracted elementwise with dK/dtheta.  All dense products run on the 128x128 DMMA core through one
// general kernel (k_tgemm) over rectangular tile matrices; the two M x M factorisations reuse gpr2.cuh.
#pragma once
#include "gpr2.cuh"

namespace gpsat {

constexpr double SGPR_JITTER = 1e-6;   // gpflow.config.default_jitter()

// rectangular tile matrix: tile (i, j) of slot s at base + s*slot_stride + (i*ld + j)*4096
struct TMat {
  double* base;
  long slot_stride;
  int ld;
  __device__ __forceinline__ double* tile(int s, int i, int j) const {
    return base + (long)s * slot_stride + ((long)i * ld + j) * TILE_ELEMS;
  }
};

struct SgCtx {
  int S, D, kid, mbmax, nbnmax;      // M-blocks / N-blocks (64) maxima over the batch
  const int* slot_expert;            // [S] expert held by the slot (cz pool's bookkeeping)
  const int* active;                 // [S]
  const double* theta;               // [S][MAXP] (cz.theta)
  const double* zcoords;             // cz.coords [S][MAXD][npmax_z] (already / coords_scale)
  int npmax_z;
  const int* mz;                     // cz.n [S]: M per slot
  // data CSR (raw units)
  const double* xcoords;             // [sumN][D]
  const double* yobs;                // [sumN]
  const long long* xoff;             // [E+1]
  double coords_scale[MAXD];
  double obs_scale;
  const double* ymean;               // [E] mean subtracted from y (zeros unless obs_mean = 'local')
  // per-slot work arrays
  TMat XLF, BF, EF, WF;              // M x M (mbmax x mbmax tiles)
  TMat KUF, AP;                      // M x N (mbmax x nbnmax tiles)
  double* vec;                       // [S][8][vlen]: ay, u, w (length M pad) ; s, r, yn (length N pad)
  int vlen;
  double* scal;                      // [S][16] scalars
  double* gpart;                     // [S][gp_n][NG]
  int gp_n;
  int* mb;                           // [S] ceil(M/64)
  int* mb1;                          // [S] M/64 + 1 (tiles covering the augmented row M)
  int* nbn;                          // [S] ceil(N/64)
  int* npb;                          // [S] ceil(P/64) (prediction)
  double* beta;                      // [S]
};
enum { V_AY = 0, V_U = 1, V_W = 2, V_S = 3, V_R = 4, V_Y = 5, V_TRB = 6, V_TRBINV = 7 };
enum { SC_YY = 0, SC_AYU = 3, SC_UU = 4 };

__device__ __forceinline__ double* sg_vec(const SgCtx& g, int s, int which) {
  return g.vec + ((long)s * 8 + which) * g.vlen;
}

// ---- per-round setup: block counts, beta, normalised y, y'y ----  grid (S), 256 threads
__global__ void __launch_bounds__(256) k_sg_setup(SgCtx g) {
  __shared__ double red[8];
  const int s = blockIdx.x;
  if (!g.active[s]) return;
  const int e = g.slot_expert[s];
  const long long o0 = g.xoff[e];
  const int N = (int)(g.xoff[e + 1] - o0), M = g.mz[s];
  double* yn = sg_vec(g, s, V_Y);
  const double mu = g.ymean[e];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < g.nbnmax * TB; i += 256) {
    const double y = (i < N) ? (g.yobs[o0 + i] - mu) / g.obs_scale : 0.0;
    yn[i] = y;
    v[0] += y * y;
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) {
    g.mb[s] = (M + TB - 1) / TB;
    g.mb1[s] = M / TB + 1;
    g.nbn[s] = (N + TB - 1) / TB;
    g.beta[s] = 1.0 / g.theta[s * MAXP + g.D + 1];
    g.scal[s * 16 + SC_YY] = v[0];
  }
}

// ---- Kuf tiles: K(Z block i, X block j), zero outside M x N.  grid (nbnmax, mbmax, S), 256 threads ----
__global__ void __launch_bounds__(256, 4) k_sg_build_uf(SgCtx g, TMat out, const double* pcoords,
                                                        const long long* poff) {
  __shared__ double xz[MAXD * TB], xx[MAXD * TB];
  const int s = blockIdx.z;
  if (!g.active[s]) return;
  const int e = g.slot_expert[s];
  const double* xc = pcoords ? pcoords : g.xcoords;
  const long long* off = poff ? poff : g.xoff;
  const long long o0 = off[e];
  const int N = (int)(off[e + 1] - o0), M = g.mz[s];
  const int i = blockIdx.y, j = blockIdx.x;
  if (i * TB >= M || j * TB >= N) return;
  const double* th = g.theta + s * MAXP;
  stage_coords(xz, g.zcoords + (long)s * MAXD * g.npmax_z, g.npmax_z, g.D, th, i, M);
  for (int t = threadIdx.x; t < g.D * TB; t += 256) {
    const int d = t / TB, m = t % TB, gx = j * TB + m;
    xx[d * TB + m] = (gx < N) ? xc[(o0 + gx) * g.D + d] / (g.coords_scale[d] * th[d]) : 0.0;
  }
  __syncthreads();
  const double kvar = th[g.D];
  const int m = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, gi = i * TB + m;
  double* o = out.tile(s, i, j);
  double xm[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) xm[d] = (d < g.D) ? xz[d * TB + m] : 0.0;
#pragma unroll
  for (int cc = 0; cc < 16; cc += 2) {
    double v[2];
#pragma unroll
    for (int e2 = 0; e2 < 2; ++e2) {
      const int n = c0 + cc + e2;
      double val = 0.0;
      if (gi < M && j * TB + n < N) {
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
          if (d < g.D) {
            const double df = xm[d] - xx[d * TB + n];
            r2 += df * df;
          }
        val = kern_value(g.kid, r2, kvar);
      }
      v[e2] = val;
    }
    *reinterpret_cast<double2*>(o + swz(m, c0 + cc)) = make_double2(v[0], v[1]);
  }
}

// ---- general batched tile GEMM:  C(I,J) = alpha * sum_k op(A)(I,k) op(B)(k,J) [+ I] ----
//   TA  = false: op(A)(i,k) = A.tile(i,k)      TA  = true: op(A)(i,k) = A.tile(k,i)^T
//   TBm = false: op(B)(k,j) = B.tile(j,k)^T    TBm = true: op(B)(k,j) = B.tile(k,j)
//   kmode: 0 all k; 1 k <= row tile; 2 k >= row tile; 3 k >= column tile   (triangular operands)
//   flags: 1 add identity on the diagonal; 2 lower supertiles only; 4 alpha = beta[s]; 8 alpha = -1
struct TGemm {
  TMat A, B, C;
  const int* mt;   // [S] output row tiles
  const int* nt;   // [S] output column tiles
  const int* kt;   // [S] inner tiles
  const int* active;
  const double* beta;
  int kmode, flags, diag_limit_from_mz;
  const int* mz;   // rows/cols >= mz[s] get no identity
};

template <bool TA, bool TBm>
__global__ void __launch_bounds__(NTHREADS, 1) k_tgemm(TGemm p, int njs) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  if (!p.active[s]) return;
  const int I = blockIdx.x / njs, J = blockIdx.x % njs;
  const int mt = p.mt[s], nt = p.nt[s], kt = p.kt[s];
  if (2 * I >= mt || 2 * J >= nt) return;
  if ((p.flags & 2) && J > I) return;
  int kbeg = 0, kend = kt;
  if (p.kmode == 1) kend = min(kt, 2 * I + 2);
  if (p.kmode == 2) kbeg = 2 * I;
  if (p.kmode == 3) kbeg = 2 * J;
  Frag2 f;
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  gemm2_pipeline<TA, TBm>(
      acc, smem, pipe, kbeg, kend,
      [&](int k, int t) -> const double* {
        const int i = 2 * I + t;
        if (i >= mt) return nullptr;
        if (p.kmode == 1 && k > i) return nullptr;
        if (p.kmode == 2 && k < i) return nullptr;
        return TA ? p.A.tile(s, k, i) : p.A.tile(s, i, k);
      },
      [&](int k, int t) -> const double* {
        const int j = 2 * J + t;
        if (j >= nt) return nullptr;
        if (p.kmode == 3 && k < j) return nullptr;
        return TBm ? p.B.tile(s, k, j) : p.B.tile(s, j, k);
      },
      f);
  const int ti = 2 * I + f.ta, tj = 2 * J + f.tb;
  if (ti >= mt || tj >= nt) return;
  if ((p.flags & 2) && tj > ti) return;
  double alpha = 1.0;
  if (p.flags & 4) alpha = p.beta[s];
  if (p.flags & 8) alpha = -1.0;
  double* ct = p.C.tile(s, ti, tj);
  const int M = (p.flags & 1) ? p.mz[s] : 0;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int m = f.row(mi);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int n = f.col(ni);
      double v0 = alpha * acc.c[mi][ni][0], v1 = alpha * acc.c[mi][ni][1];
      if ((p.flags & 1) && ti == tj) {
        if (m == n && ti * TB + m < M) v0 += 1.0;
        if (m == n + 1 && ti * TB + m < M) v1 += 1.0;
      }
      *reinterpret_cast<double2*>(ct + swz(m, n)) = make_double2(v0, v1);
    }
  }
}

// ---- unpack X_L = L^-1 (cz pool, packed / x_tile) into a rectangular M x M tile matrix (zero rows/cols >= M,
//      explicit zero tiles above the diagonal).  which = 1: same for the B pool INCLUDING its augmented row
//      (rows <= M kept).  grid (mbmax+1, mbmax+1, S) ----
__global__ void __launch_bounds__(256) k_sg_unpack_x(SlotCtx c, SgCtx g, TMat out, int keep_aug) {
  const int s = blockIdx.z;
  if (!g.active[s]) return;
  const int M = c.n[s], nb = c.nb[s];
  const int i = blockIdx.y, j = blockIdx.x;
  const int lim = keep_aug ? M + 1 : M;
  if (i * TB >= lim || j * TB >= lim || i >= nb || j >= nb) return;
  double* o = out.tile(s, i, j);
  const double* src = (j <= i) ? x_tile(c, s, i, j) : nullptr;
  for (int t = threadIdx.x; t < TILE_ELEMS; t += 256) {
    const int r = t >> 6, cc = t & 63;
    double v = 0.0;
    if (src && i * TB + r < lim && j * TB + cc < M) v = src[swz(r, cc)];
    o[swz(r, cc)] = v;
  }
}

// ---- A'y (M) and the SYRK result -> packed K_aug tiles of the B pool (+ augmented row beta A'y) ----
// grid (mbmax, S): block row i
__global__ void __launch_bounds__(256) k_sg_ay(SgCtx g) {
  __shared__ double part[4][TB];
  const int s = blockIdx.y, i = blockIdx.x;
  if (!g.active[s] || i >= g.mb[s]) return;
  const int nbn = g.nbn[s];
  const double* yn = sg_vec(g, s, V_Y);
  const int m = threadIdx.x & 63, q = threadIdx.x >> 6;
  double acc = 0.0;
  for (int j = q; j < nbn; j += 4) {
    const double* t = g.AP.tile(s, i, j);
    for (int cc = 0; cc < TB; ++cc) acc += t[swz(m, cc)] * yn[j * TB + cc];
  }
  part[q][m] = acc;
  __syncthreads();
  if (threadIdx.x < TB) sg_vec(g, s, V_AY)[i * TB + m] = (part[0][m] + part[1][m]) + (part[2][m] + part[3][m]);
}

// grid (nb_b (nb_b+1)/2 packed tiles of the B pool, S)
__global__ void __launch_bounds__(256) k_sg_pack_b(SlotCtx cb, SgCtx g) {
  const int s = blockIdx.y;
  if (!g.active[s]) return;
  int i, j;
  tri_decode(blockIdx.x, i, j);
  const int nb = cb.nb[s], M = cb.n[s];
  if (i >= nb) return;
  double* o = cb.Kt + (long)s * cb.tile_stride + tri_index(i, j) * TILE_ELEMS;
  const int mb = g.mb[s];
  const double* src = (i < mb && j < mb) ? g.BF.tile(s, i, j) : nullptr;
  const double* ay = sg_vec(g, s, V_AY);
  const double beta = g.beta[s];
  __shared__ double red[8];
  double tr[1] = {0.0};
  for (int t = threadIdx.x; t < TILE_ELEMS; t += 256) {
    const int r = t >> 6, cc = t & 63, gi = i * TB + r, gj = j * TB + cc;
    double v;
    if (gi < M && gj < M) {
      v = src[swz(r, cc)];
      if (gi == gj) tr[0] += v;
    } else if (gi == M && gj < M) v = beta * ay[gj];
    else v = (gi == gj) ? 1.0 : 0.0;
    o[swz(r, cc)] = v;
  }
  if (i == j && i < mb) {        // tr(B) per diagonal tile (deterministic partials, summed in k_sg_finalize)
    block_sum<1>(tr, red);
    if (threadIdx.x == 0) sg_vec(g, s, V_TRB)[i] = tr[0];
  }
}

// ---- after the B pool's potrf + trtri + lauum:  u, E = beta (I - B^-1), W = 2I - B^-1 - B (full, symmetric),
//      traces.  grid (mbmax, mbmax, S) ----
__global__ void __launch_bounds__(256) k_sg_prep(SlotCtx cb, SgCtx g) {
  __shared__ double red[2 * 8];
  const int s = blockIdx.z;
  if (!g.active[s]) return;
  const int i = blockIdx.y, j = blockIdx.x, mb = g.mb[s];
  if (i >= mb || j >= mb) return;
  const int M = cb.n[s], nb = cb.nb[s];
  const int bN = nb - 1, rN = M - bN * TB;
  const double beta = g.beta[s];
  // X_B'X_B = B^-1 + (beta u)(beta u)' ; the augmented row of X_B is -(beta u)'
  const bool lower = (i >= j);
  const double* kt = cb.Kt + (long)s * cb.tile_stride + (lower ? tri_index(i, j) : tri_index(j, i)) * TILE_ELEMS;
  const double* bt = lower ? g.BF.tile(s, i, j) : g.BF.tile(s, j, i);
  const double* xi = x_tile(cb, s, bN, i);
  const double* xj = x_tile(cb, s, bN, j);
  double* et = g.EF.tile(s, i, j);
  double* wt = g.WF.tile(s, i, j);
  double v[2] = {0.0, 0.0};
  for (int t = threadIdx.x; t < TILE_ELEMS; t += 256) {
    const int r = t >> 6, cc = t & 63, gi = i * TB + r, gj = j * TB + cc;
    double e = 0.0, w = 0.0;
    if (gi < M && gj < M) {
      const int rr = lower ? r : cc, c2 = lower ? cc : r;
      const double bu_i = -xi[swz(rN, r)], bu_j = -xj[swz(rN, cc)];
      const double binv = kt[swz(rr, c2)] - bu_i * bu_j;
      const double b = bt[swz(rr, c2)];
      const double id = (gi == gj) ? 1.0 : 0.0;
      e = beta * (id - binv);
      w = 2.0 * id - binv - b;
      if (gi == gj) v[1] += binv;
    }
    et[swz(r, cc)] = e;
    wt[swz(r, cc)] = w;
  }
  if (i == j) {
    block_sum<2>(v, red);
    if (threadIdx.x == 0) sg_vec(g, s, V_TRBINV)[i] = v[1];
  }
  if (j == 0 && threadIdx.x < TB) {
    const int gi = i * TB + threadIdx.x;
    sg_vec(g, s, V_U)[gi] = (gi < M) ? -xi[swz(rN, threadIdx.x)] / beta : 0.0;
  }
}

// ---- w = X_L' u (M), s = A'^T u (N), r = y - beta s, and the scalars (A'y)'u, u'u.  grid (mbmax + nbnmax, S) ----
__global__ void __launch_bounds__(256) k_sg_vec2(SgCtx g) {
  __shared__ double part[4][TB];
  __shared__ double red[2 * 8];
  const int s = blockIdx.y;
  if (!g.active[s]) return;
  const int mb = g.mb[s], nbn = g.nbn[s];
  const double* u = sg_vec(g, s, V_U);
  const int m = threadIdx.x & 63, q = threadIdx.x >> 6;
  if ((int)blockIdx.x < g.mbmax) {
    const int i = blockIdx.x;
    if (i >= mb) return;
    // w_i = sum_k X_L(k, i)' u_k   (column i of X_L: tiles (k, i), k >= i)
    double acc = 0.0;
    for (int k = i + q; k < mb; k += 4) {
      const double* t = g.XLF.tile(s, k, i);
      for (int rr = 0; rr < TB; ++rr) acc += t[swz(rr, m)] * u[k * TB + rr];
    }
    part[q][m] = acc;
    __syncthreads();
    if (threadIdx.x < TB) sg_vec(g, s, V_W)[i * TB + m] = (part[0][m] + part[1][m]) + (part[2][m] + part[3][m]);
    if (i == 0) {
      const double* ay = sg_vec(g, s, V_AY);
      double v[2] = {0.0, 0.0};
      for (int k = threadIdx.x; k < mb * TB; k += 256) {
        v[0] += ay[k] * u[k];
        v[1] += u[k] * u[k];
      }
      block_sum<2>(v, red);
      if (threadIdx.x == 0) {
        g.scal[s * 16 + SC_AYU] = v[0];
        g.scal[s * 16 + SC_UU] = v[1];
      }
    }
  } else {
    const int j = blockIdx.x - g.mbmax;
    if (j >= nbn) return;
    double acc = 0.0;
    for (int k = q; k < mb; k += 4) {
      const double* t = g.AP.tile(s, k, j);
      for (int rr = 0; rr < TB; ++rr) acc += t[swz(rr, m)] * u[k * TB + rr];
    }
    part[q][m] = acc;
    __syncthreads();
    if (threadIdx.x < TB) {
      const double sv = (part[0][m] + part[1][m]) + (part[2][m] + part[3][m]);
      sg_vec(g, s, V_S)[j * TB + m] = sv;
      sg_vec(g, s, V_R)[j * TB + m] = sg_vec(g, s, V_Y)[j * TB + m] - g.beta[s] * sv;
    }
  }
}

// ---- gradient contractions.  uf: grid (nbnmax, mbmax, S) over Kuf tiles with G = GUF + beta^2 w r';
//      uu: grid (mbmax, mbmax, S) over Kuu tiles with G = 1/2 G1 - 1/2 beta^2 w w'.  256 threads ----
template <bool UU>
__global__ void __launch_bounds__(256, 3) k_sg_trace(SgCtx g, TMat G) {
  __shared__ double xa[MAXD * TB], xb[MAXD * TB], wa[TB], rb[TB], red[NG * 8];
  const int s = blockIdx.z;
  if (!g.active[s]) return;
  const int e = g.slot_expert[s];
  const int i = blockIdx.y, j = blockIdx.x;
  const int M = g.mz[s];
  const long long o0 = g.xoff[e];
  const int N = UU ? M : (int)(g.xoff[e + 1] - o0);
  if (i * TB >= M || j * TB >= N) return;
  const double* th = g.theta + s * MAXP;
  stage_coords(xa, g.zcoords + (long)s * MAXD * g.npmax_z, g.npmax_z, g.D, th, i, M);
  if (UU) {
    stage_coords(xb, g.zcoords + (long)s * MAXD * g.npmax_z, g.npmax_z, g.D, th, j, M);
  } else {
    for (int t = threadIdx.x; t < g.D * TB; t += 256) {
      const int d = t / TB, m = t % TB, gx = j * TB + m;
      xb[d * TB + m] = (gx < N) ? g.xcoords[(o0 + gx) * g.D + d] / (g.coords_scale[d] * th[d]) : 0.0;
    }
  }
  if (threadIdx.x < TB) {
    wa[threadIdx.x] = sg_vec(g, s, V_W)[i * TB + threadIdx.x];
    rb[threadIdx.x] = UU ? sg_vec(g, s, V_W)[j * TB + threadIdx.x] : sg_vec(g, s, V_R)[j * TB + threadIdx.x];
  }
  __syncthreads();
  const double kvar = th[g.D], beta = g.beta[s], b2 = beta * beta;
  const int m = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, gi = i * TB + m;
  const double* gt = G.tile(s, i, j);
  double acc[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) acc[k] = 0.0;
  if (gi < M) {
    double xm[MAXD];
#pragma unroll
    for (int d = 0; d < MAXD; ++d) xm[d] = (d < g.D) ? xa[d * TB + m] : 0.0;
    const double wm = b2 * wa[m];
#pragma unroll
    for (int cc = 0; cc < 16; cc += 2) {
      const double2 gv = *reinterpret_cast<const double2*>(gt + swz(m, c0 + cc));
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = c0 + cc + e2;
        if (j * TB + n < N) {
          const double g0 = e2 ? gv.y : gv.x;
          const double Gv = UU ? 0.5 * (g0 - wm * rb[n]) : (g0 + wm * rb[n]);
          double r2 = 0.0, d2[MAXD];
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            d2[d] = 0.0;
            if (d < g.D) {
              const double df = xm[d] - xb[d * TB + n];
              d2[d] = df * df;
              r2 += d2[d];
            }
          }
          double kv, hv;
          kern_eval(g.kid, r2, kvar, kv, hv);
          const double gh = Gv * hv;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) acc[d] += gh * d2[d];
          acc[MAXD] += Gv * kv;
        }
      }
    }
  }
  block_sum<NG>(acc, red);
  if (threadIdx.x == 0) {
    const int slot_in = UU ? (g.mbmax * g.nbnmax + i * g.mbmax + j) : (i * g.nbnmax + j);
    double* gp = g.gpart + ((long)s * g.gp_n + slot_in) * NG;
#pragma unroll
    for (int k = 0; k < NG; ++k) gp[k] = acc[k];
  }
}

// ---- -ELBO and its gradient per slot -> cz.fout / cz.gout.  grid (S) ----
__global__ void __launch_bounds__(256) k_sg_finalize(SlotCtx cz, SlotCtx cb, SgCtx g, int with_grad) {
  __shared__ double red[(NG + 1) * 8];
  const int s = blockIdx.x;
  if (!g.active[s]) return;
  const int e = g.slot_expert[s];
  const int N = (int)(g.xoff[e + 1] - g.xoff[e]), M = cz.n[s];
  const int mb = g.mb[s], nbn = g.nbn[s];
  double v[NG + 1];
#pragma unroll
  for (int k = 0; k < NG + 1; ++k) v[k] = 0.0;
  for (int k = threadIdx.x; k < cb.nb[s]; k += 256) v[NG] += cb.logdet_part[s * cb.nbmax + k];
  if (with_grad) {
    for (int t = threadIdx.x; t < mb * nbn + mb * mb; t += 256) {
      int idx;
      if (t < mb * nbn) idx = (t / nbn) * g.nbnmax + t % nbn;
      else {
        const int q = t - mb * nbn;
        idx = g.mbmax * g.nbnmax + (q / mb) * g.mbmax + q % mb;
      }
      const double* gp = g.gpart + ((long)s * g.gp_n + idx) * NG;
#pragma unroll
      for (int k = 0; k < NG; ++k) v[k] += gp[k];
    }
  }
  block_sum<NG + 1>(v, red);
  if (threadIdx.x == 0) {
    const double* th = cz.theta + s * MAXP;
    const double beta = g.beta[s], kvar = th[cz.D], nvar = th[cz.D + 1];
    const double* sc = g.scal + s * 16;
    double trB = 0.0, trBinv = 0.0;
    for (int k = 0; k < mb; ++k) {
      trB += sg_vec(g, s, V_TRB)[k];
      trBinv += sg_vec(g, s, V_TRBINV)[k];
    }
    const double trAA = (trB - M) / beta;          // tr(A'A'^T)
    double F = -0.5 * N * 1.8378770664093453 - v[NG] - 0.5 * N * log(nvar) - 0.5 * beta * sc[SC_YY] +
               0.5 * cb.quad[s] - 0.5 * beta * N * kvar + 0.5 * beta * trAA;
    if (cz.fail[s] || cb.fail[s]) F = -INFINITY;
    cz.fail[s] = 0;
    cb.fail[s] = 0;
    cz.fout[s] = -F;
    if (with_grad) {
      double* go = cz.gout + s * MAXP;
      for (int d = 0; d < cz.D; ++d) go[d] = -v[d] / th[d];
      go[cz.D] = -(v[MAXD] / kvar - 0.5 * beta * N);
      // dF/dbeta in whitened quantities (see header): tr(Sigma^-1 P) = (M - tr B^-1)/beta,
      // v'w = (A'y)'u, w'Pw = ((A'y)'u - u'u)/beta, tr(Kuu^-1 P) = tr(A'A'^T)
      const double dFdb = 0.5 * N / beta - 0.5 * (M - trBinv) / beta - 0.5 * sc[SC_YY] + beta * sc[SC_AYU] -
                          0.5 * beta * (sc[SC_AYU] - sc[SC_UU]) - 0.5 * N * kvar + 0.5 * trAA;
      go[cz.D + 1] = beta * beta * dFdb;     // d(-F)/d nvar = +beta^2 dF/dbeta
    }
  }
}

// ---- prediction epilogue: var = kvar - colsum(t1^2) + colsum(t2^2), mean = -(augmented row of X_B,aug t1).
//      t1: M x P tiles (AP), t2: (M+1) x P tiles (KUF buffer).  grid (npb, S) ----
__global__ void __launch_bounds__(256) k_sg_pred_out(SlotCtx cb, SgCtx g, TMat T1, TMat T2, const long long* poff,
                                                     double* fmean, double* fvar, double* yvar) {
  __shared__ double p1[4][TB], p2[4][TB];
  const int s = blockIdx.y, j = blockIdx.x;
  if (!g.active[s]) return;
  const int e = g.slot_expert[s];
  const long long o0 = poff[e];
  const int P = (int)(poff[e + 1] - o0), M = cb.n[s];
  if (j * TB >= P) return;
  const int mb = g.mb[s], mb1 = M / TB + 1;
  const int n = threadIdx.x & 63, q = threadIdx.x >> 6;
  double a1 = 0.0, a2 = 0.0;
  for (int k = q; k < mb1; k += 4) {
    const double* t2 = T2.tile(s, k, j);
    const double* t1 = (k < mb) ? T1.tile(s, k, j) : nullptr;
    for (int rr = 0; rr < TB; ++rr) {
      if (k * TB + rr < M) {
        const double b = t2[swz(rr, n)];
        a2 += b * b;
        if (t1) {
          const double a = t1[swz(rr, n)];
          a1 += a * a;
        }
      }
    }
  }
  p1[q][n] = a1;
  p2[q][n] = a2;
  __syncthreads();
  if (threadIdx.x < TB && j * TB + n < P) {
    const double* th = g.theta + s * MAXP;
    const double s1 = (p1[0][n] + p1[1][n]) + (p1[2][n] + p1[3][n]);
    const double s2 = (p2[0][n] + p2[1][n]) + (p2[2][n] + p2[3][n]);
    const double var = th[g.D] - s1 + s2;
    const int bM = M / TB, rM = M - bM * TB;
    const long long o = o0 + j * TB + n;
    fmean[o] = -T2.tile(s, bM, j)[swz(rM, n)];
    fvar[o] = var;
    yvar[o] = var + th[g.D + 1];
  }
}

}  // namespace gpsat
