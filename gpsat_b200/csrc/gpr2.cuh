// gpsat_b200: batched exact-GPR kernels, generation 2 (128x128 supertile DMMA core).
//
// One objective evaluation for every active slot is
//   k_build         K_aug tiles (kernel matrix + noise, augmented with the observation row) -> Kt
//                   FP64 elementwise, high occupancy (the DMMA and DFMA pipes are the same hardware
//                   on sm_100, so elementwise work cannot hide behind tensor work -- only its latency
//                   can be hidden, by running it in its own many-warps kernel)
//   k_potrf_*2      left-looking blocked Cholesky with 128-wide panels; the diagonal CTA factorises
//                   its 128x128 block in shared memory and also emits the block's inverse
//   k_quad          a'a from the augmented row of the factor (row N of L_aug is a' = (L^-1 y)')
//   k_trtri_*       X = L_aug^-1 by recursive doubling: at level l every 2^l-supertile block gets its
//                   lower-left square X21 = -X22 (L21 X11) from two batched GEMM passes.  L21 is dead
//                   after pass 1, so X21 is written over it: the inverse ends up with its 128x128
//                   diagonal blocks in Xt and everything else in Lt (see x_tile()).
//   k_lauum2        tiles of X'X = K_y^-1 + alpha alpha' -> Kt (K is dead after the factorisation)
//   k_grad_trace    G_k = sum_ij (K^-1 - alpha alpha')_ij dK_ij/dtheta_k, FP64 elementwise, high occupancy
//   k_finalize2     -LML and d(-LML)/dtheta
// Prediction (row F1): k_build_xp writes cross-covariance tiles, k_predict2 accumulates
// A = X K_xp supertile by supertile with a fused column sum of squares; the augmented row of X
// (-alpha') yields the posterior mean.
#pragma once
#include "gemm_core.cuh"
#include "gemm2.cuh"

namespace gpsat {

constexpr int NG = MAXP;                              // gradient partials per tile
// diagonal-block workspace (diag_block_128): a[64][68] + inv[64][68] + 4 tiles
constexpr int LDA = 68;
constexpr int LDI = 68;
constexpr int DIAG_ELEMS = TB * LDA + TB * LDI + 4 * TILE_ELEMS;            // 25088 (diag_block_128 workspace)
constexpr int G2_AUX = 2 * MAXD * TB + 2 * TB + 64;
constexpr int SMEM2_ELEMS = (DIAG_ELEMS > G2_SMEM_ELEMS ? DIAG_ELEMS : G2_SMEM_ELEMS) + G2_AUX;
constexpr int SMEM2_BYTES = SMEM2_ELEMS * 8;

struct SlotCtx {
  int S, D, kid, nbmax, npmax, ntmax;
  double nvar_override;     // >= 0: diagonal term of k_build instead of theta's likelihood variance (SGPR jitter)
  long tile_stride;         // doubles per slot in Lt / Xt / Kt (= ntmax * TILE_ELEMS)
  double* Lt;               // [S][ntmax][4096] packed lower tiles: L_aug, later the off-diagonal part of X
  double* Xt;               // [S][ntmax][4096] diagonal 128-blocks of X = L_aug^-1 (+ scratch T)
  double* Kt;               // [S][ntmax][4096] K_aug tiles, later X'X tiles
  double* coords;           // [S][MAXD][npmax]  coordinates / coords_scale
  double* yobs;             // [S][npmax]        (obs - mean) / scale
  int* n;                   // [S] observations per slot
  int* nb;                  // [S] 64-blocks of the augmented matrix = n/64 + 1
  int* active;              // [S]
  double* theta;            // [S][MAXP] lengthscales[D], kernel variance, likelihood variance
  double* logdet_part;      // [S][nbmax]
  double* quad;             // [S] a'a
  double* gpart;            // [S][ntmax][NG]
  int* fail;                // [S] set when a pivot is not positive
  int* pflag;               // [S] k_potrf_panel: J + 1 once the diagonal block of panel J is published (k_quad resets)
  double* fout;             // [S]  -LML
  double* gout;             // [S][MAXP] d(-LML)/dtheta (constrained parameters)
  int* timeouts;            // [1] device counter of flag-wait timeouts (handle level; reported as GPSAT_ESYNC)
};

__device__ __forceinline__ double* tile_ptr(double* base, int i, int j) {
  return base + tri_index(i, j) * TILE_ELEMS;
}
// where tile (i, j), i >= j, of X = L_aug^-1 lives after k_trtri_* (see header comment)
__device__ __forceinline__ double* x_tile(const SlotCtx& c, int s, int i, int j) {
  double* base = ((i >> 1) == (j >> 1)) ? c.Xt : c.Lt;
  return base + (long)s * c.tile_stride + tri_index(i, j) * TILE_ELEMS;
}
__device__ __forceinline__ void tri_decode(int t, int& i, int& j) {
  i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((long)(i + 1) * (i + 2) / 2 <= t) ++i;
  while ((long)i * (i + 1) / 2 > t) --i;
  j = t - (int)((long)i * (i + 1) / 2);
}

// stage the (scaled-by-1/l) coordinates of block `blk` into dst[MAXD][64]
__device__ __forceinline__ void stage_coords(double* dst, const double* coords_slot, int npmax, int D,
                                             const double* th, int blk, int N) {
  for (int t = threadIdx.x; t < D * TB; t += blockDim.x) {
    const int d = t / TB, m = t % TB, g = blk * TB + m;
    dst[d * TB + m] = (g < N) ? coords_slot[(long)d * npmax + g] / th[d] : 0.0;
  }
}

// ------------------------------------------------------------------------------------
// K1: kernel-matrix build.  grid (ntmax, S), 256 threads: thread = (row, 16-column group).
// ------------------------------------------------------------------------------------
template <int KID>
__global__ void __launch_bounds__(256, 4) k_build(SlotCtx c) {
  __shared__ double xi[MAXD * TB], xj[MAXD * TB], yj[TB];
  const int s = blockIdx.y;
  const int act = c.active[s], nb = c.nb[s], N = c.n[s];
  if (!act) return;
  int i, j;
  tri_decode(blockIdx.x, i, j);
  if (i >= nb) return;
  const double* th = c.theta + s * MAXP;
  const double* cs = c.coords + (long)s * MAXD * c.npmax;
  stage_coords(xi, cs, c.npmax, c.D, th, i, N);
  stage_coords(xj, cs, c.npmax, c.D, th, j, N);
  if (threadIdx.x < TB) {
    const int g = j * TB + threadIdx.x;
    yj[threadIdx.x] = (g < N) ? c.yobs[(long)s * c.npmax + g] : 0.0;
  }
  __syncthreads();
  const double kvar = th[c.D], nvar = (c.nvar_override >= 0.0) ? c.nvar_override : th[c.D + 1];
  const int m = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, gi = i * TB + m;
  double* out = c.Kt + (long)s * c.tile_stride + tri_index(i, j) * TILE_ELEMS;
  double xm[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) xm[d] = (d < c.D) ? xi[d * TB + m] : 0.0;
  if (i < nb - 1 && j < i) {
    // interior tile: every row and column is an observation and nothing lies on the diagonal -> no per-entry tests
    // (the loop is deliberately not unrolled: the fully unrolled body was 128 KiB of code and the warps
    //  stalled on instruction fetch -- ncu stalled_no_instruction 4.8 per issue)
#pragma unroll 1
    for (int cc = 0; cc < 16; cc += 4) {
      double v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = c0 + cc + e;
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
          if (d < c.D) {
            const double df = xm[d] - xj[d * TB + n];
            r2 += df * df;
          }
        v[e] = kern_value_t<KID>(r2, kvar);
      }
      st_global_v4(out + swz(m, c0 + cc), v[0], v[1], v[2], v[3]);   // 4 consecutive columns = one 32-byte sector
    }
    return;
  }
#pragma unroll 1
  for (int cc = 0; cc < 16; cc += 4) {
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = c0 + cc + e, gj = j * TB + n;
      double val;
      if (gi < N && gj < N) {
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
          if (d < c.D) {
            const double df = xm[d] - xj[d * TB + n];
            r2 += df * df;
          }
        val = kern_value_t<KID>(r2, kvar);
        if (gi == gj) val += nvar;
      } else if (gi == N && gj < N) {
        val = yj[n];
      } else {
        val = (gi == gj) ? 1.0 : 0.0;
      }
      v[e] = val;
    }
    st_global_v4(out + swz(m, c0 + cc), v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------
// 64x64 diagonal block: Cholesky + triangular inverse in shared memory
// a: [64][LDA] (lower part valid), inv: [64][LDI], dg: [64].  All NTHREADS (8 warps) call.
// Global indices >= N (augmented row and padding) get a forced unit pivot.
//
// Blocked by 8: warp w owns block row w.  Left-looking step jb: every warp w >= jb forms
// U = A[w][jb] - sum_kb L[w][kb] L[jb][kb]' with DMMA (register accumulator); warp jb factorises its 8x8 block
// and inverts it with every lane doing the same 8x8 arithmetic in registers (no shuffles on the critical
// chain); the warps below multiply by the inverse (U reaches the A-fragment layout by two shuffles).
// The triangular inverse is then swept by block columns, one warp per column, without block barriers:
// X[ib][jb] = -X[ib][ib] (sum_kb L[ib][kb] X[kb][jb]).   (The previous element-wise version took 60 us per
// 64x64 block and was the critical path of every Cholesky panel; this one is ~8x shorter.)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void chol8_inv(double* blk, double* invblk, double* dg8, int gbase, int N,
                                          int* fail_flag, int lane) {
  bool bad = false;
#include "chol8.inc"
  if (bad && lane == 0) *fail_flag = 1;
}

// stage probe of the diagonal-block routines: a no-op in the production kernels; the micro-benchmark (microbench.cuh)
// passes one that records clock64() per stage
struct NoProbe {
  __device__ __forceinline__ void operator()(int) const {}
};
template <class PROBE = NoProbe>
__device__ __forceinline__ void potf2_trtri_64(double* a, double* inv, double* dg, int g0, int N, int* fail_flag,
                                               PROBE probe = PROBE(), int stage0 = 0) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, q = lane >> 2, r = lane & 3;
  for (int t = tid; t < TB * LDI; t += NTHREADS) inv[t] = 0.0;
  __syncthreads();
  for (int jb = 0; jb < 8; ++jb) {
    double c0 = 0.0, c1 = 0.0;
    if (w >= jb) {
      c0 = a[(8 * w + q) * LDA + 8 * jb + 2 * r];
      c1 = a[(8 * w + q) * LDA + 8 * jb + 2 * r + 1];
      for (int kb = 0; kb < jb; ++kb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double af = -a[(8 * w + q) * LDA + 8 * kb + 4 * h + r];
          const double bf = a[(8 * jb + q) * LDA + 8 * kb + 4 * h + r];
          dmma884(c0, c1, af, bf);
        }
      }
    }
    if (w == jb) {
      a[(8 * w + q) * LDA + 8 * jb + 2 * r] = c0;
      a[(8 * w + q) * LDA + 8 * jb + 2 * r + 1] = c1;
      __syncwarp();
      chol8_inv(a + (8 * jb) * LDA + 8 * jb, inv + (8 * jb) * LDI + 8 * jb, dg + 8 * jb, g0 + 8 * jb, N, fail_flag,
                lane);
    }
    __syncthreads();   // the inverse of the diagonal 8x8 block is visible
    if (w > jb) {      // L[w][jb] = U inv(L_jj)'
      double o0 = 0.0, o1 = 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int src = q * 4 + 2 * h + (r >> 1);
        const double t0 = __shfl_sync(0xffffffffu, c0, src), t1 = __shfl_sync(0xffffffffu, c1, src);
        const double af = (r & 1) ? t1 : t0;                                   // U[q][4h + r]
        const double bf = inv[(8 * jb + q) * LDI + 8 * jb + 4 * h + r];        // B[k][n] = invL[n][k]
        dmma884(o0, o1, af, bf);
      }
      a[(8 * w + q) * LDA + 8 * jb + 2 * r] = o0;
      a[(8 * w + q) * LDA + 8 * jb + 2 * r + 1] = o1;
    }
    __syncthreads();   // block column jb of L is final
  }
  probe(stage0);
  // triangular inverse: warp w sweeps block column w
  for (int ib = w + 1; ib < 8; ++ib) {
    double t0 = 0.0, t1 = 0.0;
    for (int kb = w; kb < ib; ++kb) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double af = a[(8 * ib + q) * LDA + 8 * kb + 4 * h + r];           // L[ib][kb]
        const double bf = inv[(8 * kb + 4 * h + r) * LDI + 8 * w + q];          // X[kb][w]
        dmma884(t0, t1, af, bf);
      }
    }
    double o0 = 0.0, o1 = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double af = inv[(8 * ib + q) * LDI + 8 * ib + 4 * h + r];           // X[ib][ib]
      const int src = (4 * h + r) * 4 + (q >> 1);
      const double u0 = __shfl_sync(0xffffffffu, t0, src), u1 = __shfl_sync(0xffffffffu, t1, src);
      const double bf = (q & 1) ? u1 : u0;                                      // T[4h + r][q]
      dmma884(o0, o1, af, bf);
    }
    inv[(8 * ib + q) * LDI + 8 * w + 2 * r] = -o0;
    inv[(8 * ib + q) * LDI + 8 * w + 2 * r + 1] = -o1;
    __syncwarp();
  }
  __syncthreads();
  probe(stage0 + 1);
}

// write L (from a/dg) and X (from inv) of a factorised diagonal block: L -> gL, X -> gX and sX (all swizzled)
__device__ __forceinline__ void emit_diag(const double* a, const double* inv, const double* dg, double* gL,
                                          double* gX, double* sX) {
  for (int t = threadIdx.x; t < TILE_ELEMS; t += NTHREADS) {
    const int r = t >> 6, cc = t & 63;
    const double lv = (cc < r) ? a[r * LDA + cc] : ((cc == r) ? dg[r] : 0.0);
    const double xv = inv[r * LDI + cc];
    gL[swz(r, cc)] = lv;
    gX[swz(r, cc)] = xv;
    sX[swz(r, cc)] = xv;
  }
}
__device__ __forceinline__ void emit_logdet(const double* dg, int g0, int N, double* out) {
  if (threadIdx.x < 32) {
    double ld = 0.0;
    for (int k = threadIdx.x; k < TB; k += 32)
      if (g0 + k < N) ld += log(dg[k]);
    ld = warp_sum(ld);
    if (threadIdx.x == 0) *out = ld;
  }
}

// ------------------------------------------------------------------------------------
// 128x128 diagonal block of a panel, entirely in shared memory.  In: ws[0 ..) = C00 as a[64][LDA] (lower part),
// ws + DIAG_P1 = C10, ws + DIAG_P2 = C11 (tile images).  Out (global, tile images): L00, L10, L11 and
// X = L^-1: X00, X10, X11; log-determinant partials ld[0], ld[1].  `two` = the block has a second tile row.
// All NTHREADS threads call; ends with the global stores issued (no trailing barrier).
// ------------------------------------------------------------------------------------
constexpr int DIAG_INV = TB * LDA;
constexpr int DIAG_P0 = DIAG_INV + TB * LDI;
constexpr int DIAG_P1 = DIAG_P0 + TILE_ELEMS;
constexpr int DIAG_P2 = DIAG_P1 + TILE_ELEMS;
constexpr int DIAG_P3 = DIAG_P2 + TILE_ELEMS;
template <class PROBE = NoProbe>
__device__ __forceinline__ void diag_block_128(double* ws, double* dg, bool two, int j0, int N, int* fail_flag,
                                               double* gL00, double* gL10, double* gL11, double* gX00, double* gX10,
                                               double* gX11, double* ld, PROBE probe = PROBE()) {
  double* a = ws;
  double* inv = ws + DIAG_INV;
  double* P0 = ws + DIAG_P0;      // X00
  double* P1 = ws + DIAG_P1;      // C10, later M = L10 X00
  double* P2 = ws + DIAG_P2;      // C11, later X11
  double* P3 = ws + DIAG_P3;      // L10
  probe(0);
  potf2_trtri_64(a, inv, dg, j0 * TB, N, fail_flag, probe, 1);
  emit_diag(a, inv, dg, gL00, gX00, P0);
  emit_logdet(dg, j0 * TB, N, ld);
  __syncthreads();
  probe(3);
  if (!two) return;
  const int j1 = j0 + 1;
  FragCoord fc;
  {  // L10 = C10 X00'
    Acc t;
    t.zero();
    mma_tile<false, false>(t, P1, P0, fc);
    store_acc_swizzled(P3, t, fc);
    store_acc_swizzled(gL10, t, fc);
  }
  __syncthreads();
  probe(4);
  {  // C11' = C11 - L10 L10'
    Acc t;
    t.zero();
    mma_tile<false, false>(t, P3, P3, fc);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int m = fc.row(mi), n = fc.col(ni) + e;
          a[m * LDA + n] = P2[swz(m, n)] - t.c[mi][ni][e];
        }
  }
  potf2_trtri_64(a, inv, dg, j1 * TB, N, fail_flag, probe, 5);   // (its leading barrier closes stage 4 -> C11')
  emit_diag(a, inv, dg, gL11, gX11, P2);
  emit_logdet(dg, j1 * TB, N, ld + 1);
  __syncthreads();
  probe(7);
  {  // M = L10 X00
    Acc t;
    t.zero();
    mma_tile<false, true>(t, P3, P0, fc);
    store_acc_swizzled(P1, t, fc);
  }
  __syncthreads();
  probe(8);
  {  // X10 = -X11 M
    Acc t;
    t.zero();
    mma_tile<false, true>(t, P2, P1, fc);
    store_acc_swizzled(gX10, t, fc, -1.0);
  }
  probe(9);
}

// ------------------------------------------------------------------------------------
// Fused potrf panel J (update + triangular solve in one launch per panel: the C block never
// goes through global memory).  1-D grid, diagonal CTAs first:
//   blocks [0, S)            : slot b, supertile row I = J: update + 128x128 factorisation (diag_block_128),
//                              then publish pflag[s] = J + 1 (release)
//   blocks [S, S + S * nd)   : nd = nsr_max - J - 1; slot (b - S) / nd, row I = J + 1 + (b - S) % nd:
//                              C = K - sum_k L_Ik L_Jk' in registers -> shared memory (4 tile images), wait for the
//                              slot's flag (blocks are dispatched in index order, so its diagonal CTA started at
//                              least a wave earlier), fetch X_JJ = L_JJ^-1 (3 tiles) by TMA and form
//                              L_I,panel = C X_JJ' from shared memory.
// Shared memory: ring [0, 192 KiB) (later C [0, 128) + X10, X11 [128, 192)) + X00 [192, 224 KiB).
// ------------------------------------------------------------------------------------
constexpr int PANEL_SMEM_ELEMS = G2_SMEM_ELEMS + TILE_ELEMS + 64;   // + X00 + dg of the diagonal CTA
constexpr int PANEL_SMEM_BYTES = PANEL_SMEM_ELEMS * 8;
static_assert(DIAG_ELEMS + 64 <= PANEL_SMEM_ELEMS, "diagonal workspace must fit");
static_assert(PANEL_SMEM_BYTES + 256 <= 232448, "over the 227 KiB shared-memory limit");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// block0: index of this launch's first block in the panel's block numbering.  One launch covers the whole panel
// (block0 = 0); the safe mode of the host (api.cu: after a flag-wait timeout, or GPSAT_SAFE_PANEL=1) launches the
// diagonal blocks [0, S) and the off-diagonal blocks [S, ...) as two kernels, so that the flag is set before any waiter
// exists and nothing depends on the order in which the hardware dispatches CTAs.
//
// la = 1, LOOK-AHEAD mode (api.cu: small matrices): the launch holds the off-diagonal blocks of panel J only
// (block0 = S), numbered with row I = J + 1 of every slot first.  That CTA owns ALL of row J + 1 of the factor once its
// own L_{J+1, panel J} is written, so it carries on as the diagonal CTA of panel J + 1 (update over k < 2J + 2,
// factorisation, flag = J + 2) inside the same launch: the diagonal block of every panel is published a whole launch
// before its consumers exist, no CTA ever spins on a flag while holding an SM, and panel J + 1's 41 us serial
// factorisation overlaps panel J's remaining off-diagonal CTAs.  Panel 0's diagonal blocks are their own launch
// (block0 = 0, S blocks, la = 0).  Same arithmetic in the same order: results are bit-identical to the fused launch.
__global__ void __launch_bounds__(NTHREADS, 1) k_potrf_panel(SlotCtx c, int J, int nsr_max, int block0, int la) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) uint64_t xbar[2];
  int s, I;
  const int bid = (int)blockIdx.x + block0;
  if (bid < c.S) {
    s = bid;
    I = J;
  } else if (la) {
    const int b = bid - c.S, nd = nsr_max - J - 1;
    if (b < c.S) {               // the look-ahead CTAs (row J + 1) first: they are the long ones
      s = b;
      I = J + 1;
    } else {
      s = (b - c.S) / (nd - 1);
      I = J + 2 + (b - c.S) % (nd - 1);
    }
  } else {
    const int b = bid - c.S, nd = nsr_max - J - 1;
    s = b / nd;
    I = J + 1 + b % nd;
  }
  // the three words of slot state are fetched together: one L2 round trip instead of three before the first copy
  const int act = c.active[s], nb = c.nb[s], N = c.n[s];
  if (!act) return;
  int Jc = J;                                 // the panel this CTA is working on (J, then J + 1 in look-ahead mode)
  int j0 = 2 * J, j1 = 2 * J + 1;
  if (2 * I >= nb) return;
  double* Lt = c.Lt + (long)s * c.tile_stride;
  double* Kt = c.Kt + (long)s * c.tile_stride;
  double* Xt = c.Xt + (long)s * c.tile_stride;
  bool two = (j1 < nb);
  G2Pipe pipe;
  if (threadIdx.x == 0) {
    mbar_init(xbar, 1);
    mbar_init(xbar + 1, 1);
  }
  pipe.init();   // fences the mbarrier inits and syncs
  auto a_of = [&](int k, int t) -> const double* { return (2 * I + t < nb) ? tile_ptr(Lt, 2 * I + t, k) : nullptr; };
  auto b_of = [&](int k, int t) -> const double* { return (j0 + t < nb) ? tile_ptr(Lt, j0 + t, k) : nullptr; };
  auto k_of = [&](int e, int t) -> const double* {
    const int a = 2 * I + e, b = j0 + t;
    return (a < nb && b < nb && b <= a) ? tile_ptr(Kt, a, b) : nullptr;
  };
  if (I != Jc) {
    // ---- off-diagonal supertile: C = K - sum_k L_Ik L_Jk', then L_I,panel = C X_JJ' ----
    // (rows 32-63 of the last tile row are skipped when they are all padding: their C and L entries are exact zeros)
    Frag2H f(2 * I, nb - 1, (N - (nb - 1) * TB) < 32);
    // L_I,panel = [C0 C1] X_JJ' with X_JJ = [[X00, 0], [X10, X11]]:  column j0 = C0 X00',  column j1 = C0 X10' + C1 X11'.
    // Step 1 (k = j1, column-j1 warps only) needs X11 alone: it is fetched early into the extra tile behind the ring,
    // and X00 / X10 for step 2 land in the third ring slot while step 1 runs.
    double* sC = smem;                          // 4 tile images (ta, tb) at (2 * ta + tb) * TILE_ELEMS
    double* sX1 = smem + 4 * TILE_ELEMS;        // X00, X10 (two tile columns) -- third ring slot
    double* sXe = smem + G2_SMEM_ELEMS;         // X11 (or X00 when the panel has one tile column) -- outside the ring
    const double* xe_src = two ? tile_ptr(Xt, j1, j1) : tile_ptr(Xt, j0, j0);
    bool xe_issued = false;
    // (thread 32 owns the X_JJ fetches, so that thread 0 goes straight to the first ring copies)
    // (the flag only grows within an evaluation -- a look-ahead CTA may already have published panel J + 1 -- hence >=)
    if (threadIdx.x == 32 && ld_acquire_gpu(c.pflag + s) >= J + 1) {   // already published: the common case
      fence_proxy_async_all();
      mbar_expect_tx(xbar, TILE_BYTES);
      bulk_g2s(sXe, xe_src, TILE_BYTES, xbar);
      xe_issued = true;
    }
    Acc2 acc;
    acc.zero();
    const int ti = 2 * I + f.ta, tj = j0 + f.tb;
    const bool valid = (ti < nb) && (tj < nb) && (tj <= ti);
    const double* ktile[2];
    gemm2_pipeline_t<false, false, true>(acc, smem, pipe, 0, j0, a_of, b_of, f, k_of, ktile);
    if (valid) {
      const double* kt = ktile[f.ta] + f.tb * TILE_ELEMS;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const double2 kv = *reinterpret_cast<const double2*>(kt + swz(f.row(mi), f.col(ni)));
          acc.c[mi][ni][0] = kv.x - acc.c[mi][ni][0];
          acc.c[mi][ni][1] = kv.y - acc.c[mi][ni][1];
        }
    }
    __syncthreads();   // every warp has read its K tile: the ring is free
    if (threadIdx.x == 32) {
      // Blocks are dispatched in index order, so the slot's diagonal CTA (index < S) is resident or finished by
      // the time this one runs and the wait is short.  It is bounded anyway (~1 s): a lost flag marks the slot
      // as failed (objective = +inf) and is reported to the caller as GPSAT_ESYNC instead of hanging the device.
      int spins = 0;
      while (ld_acquire_gpu(c.pflag + s) < J + 1) {
        __nanosleep(200);
        if (++spins > 4000000) {
          c.fail[s] = 1;                                  // the evaluation is unusable: f = +inf for this slot ...
          if (c.timeouts) atomicAdd(c.timeouts, 1);       // ... and the call reports GPSAT_ESYNC, not "non-PD"
          break;
        }
      }
      fence_proxy_async_all();
      if (!xe_issued) {
        mbar_expect_tx(xbar, TILE_BYTES);
        bulk_g2s(sXe, xe_src, TILE_BYTES, xbar);
      }
      if (two) {
        mbar_expect_tx(xbar + 1, 2 * TILE_BYTES);
        bulk_g2s(sX1, tile_ptr(Xt, j0, j0), TILE_BYTES, xbar + 1);
        bulk_g2s(sX1 + TILE_ELEMS, tile_ptr(Xt, j1, j0), TILE_BYTES, xbar + 1);
      }
    }
    if (ti < nb) store_acc2(sC + (2 * f.ta + f.tb) * TILE_ELEMS, acc, f);
    __syncthreads();
    acc.zero();
    mbar_wait(xbar, 0);
    if (ti < nb && f.tb == (two ? 1 : 0)) {         // step 1: C1 X11'  (one tile column: C0 X00')
      const double* As = sC + (2 * f.ta + (two ? 1 : 0)) * TILE_ELEMS;
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) mma_half<false, false>(acc, As + kh * HALF_ELEMS, sXe + kh * HALF_ELEMS, f);
    }
    if (two) {
      mbar_wait(xbar + 1, 0);
      if (ti < nb) {                                // step 2: C0 X00' -> column j0, C0 X10' -> column j1
        const double* As = sC + (2 * f.ta) * TILE_ELEMS;
        const double* Bs = sX1 + f.tb * TILE_ELEMS;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) mma_half<false, false>(acc, As + kh * HALF_ELEMS, Bs + kh * HALF_ELEMS, f);
      }
    }
    if (ti < nb && tj < nb) store_acc2(tile_ptr(Lt, ti, tj), acc, f);
    if (!(la && I == J + 1)) return;
    // ---- look-ahead: this CTA now holds every tile of row J + 1 of the factor; become the diagonal CTA of panel J + 1.
    // Its own stores of L_{J+1, panel J} are read back by TMA below: gpu-scope fence for the global writes, proxy fence
    // generic -> async for them and for the shared memory the ring is about to overwrite, then the block barrier.
    __threadfence();
    fence_proxy_async_all();
    __syncthreads();
    Jc = J + 1;
    j0 = 2 * Jc;
    j1 = j0 + 1;
    two = (j1 < nb);
  }
  // ---- diagonal 128x128 block: update of the three lower tiles with the balanced diagonal warp map (gemm2.cuh:
  //      a k-step costs 3/4 of a full supertile's), then the in-CTA factorisation ----
  Frag2D f;
  Acc2 acc;
  acc.zero();
  const int ti = 2 * I + f.ta, tj = j0 + f.tb;
  const bool valid = (ti < nb) && (tj < nb);
  const double* ktile[2];
  gemm2_pipeline_t<false, false, true>(acc, smem, pipe, 0, j0, a_of, b_of, f, k_of, ktile);
  if (valid) {
    const double* kt = ktile[f.ta] + f.tb * TILE_ELEMS;
    const int m0 = f.mi_begin(), m1 = f.mi_end();
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
      if (mi < m0 || mi >= m1) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const double2 kv = *reinterpret_cast<const double2*>(kt + swz(f.row(mi), f.col(ni)));
        acc.c[mi][ni][0] = kv.x - acc.c[mi][ni][0];
        acc.c[mi][ni][1] = kv.y - acc.c[mi][ni][1];
      }
    }
  }
  __syncthreads();   // every warp has read its K tile: the ring is free
  double* dg = smem + G2_SMEM_ELEMS + TILE_ELEMS;
  if (f.ta == 0) {                        // tile (0,0): warps 0 and 1 hold all 64 rows of their slabs
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        smem[f.row(mi) * LDA + f.col(ni)] = acc.c[mi][ni][0];
        smem[f.row(mi) * LDA + f.col(ni) + 1] = acc.c[mi][ni][1];
      }
  } else if (two && f.tb == 0) {
    store_acc2(smem + DIAG_P1, acc, f);
  } else if (two) {
    store_acc2(smem + DIAG_P2, acc, f);   // (1,1): each of the four warps stores its 32 rows
  }
  diag_block_128(smem, dg, two, j0, N, c.fail + s, tile_ptr(Lt, j0, j0), two ? tile_ptr(Lt, j1, j0) : nullptr,
                 two ? tile_ptr(Lt, j1, j1) : nullptr, tile_ptr(Xt, j0, j0), two ? tile_ptr(Xt, j1, j0) : nullptr,
                 two ? tile_ptr(Xt, j1, j1) : nullptr, c.logdet_part + s * c.nbmax + j0);
  // publish: every thread's global stores -> gpu scope, then one release store of the flag
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) st_release_gpu(c.pflag + s, Jc + 1);
}

// a'a from the augmented row of the factor (must run before k_trtri_* recycles Lt).  grid (S)
__global__ void __launch_bounds__(NTHREADS) k_quad(SlotCtx c) {
  __shared__ double red[NTHREADS / 32];
  const int s = blockIdx.x;
  if (!c.active[s]) return;
  const int N = c.n[s], nb = c.nb[s];
  const int bN = nb - 1, rN = N - bN * TB;
  const double* Lt = c.Lt + (long)s * c.tile_stride;
  double v[1] = {0.0};
  for (int idx = threadIdx.x; idx < N; idx += NTHREADS) {
    const double a = Lt[tri_index(bN, idx >> 6) * TILE_ELEMS + swz(rN, idx & 63)];
    v[0] += a * a;
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) {
    c.quad[s] = v[0];
    c.pflag[s] = 0;   // the factorisation of this evaluation is complete: re-arm the panel flags
  }
}

// ------------------------------------------------------------------------------------
// trtri level (block = B supertiles, half h = B/2): lower-left square of every block.
//   pass 1:  T   = L21 X11   -> Xt (scratch)          pass 2:  X21 = -X22 T  -> Lt (over the dead L21)
// grid (nblk * h * h, S)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool trtri_decode(int bx, int h, int nsr, int& P, int& Q, int& mid) {
  const int hh = h * h;
  const int m = bx / hh, rem = bx - m * hh;
  const int s0 = m * 2 * h;
  mid = s0 + h;
  P = mid + rem / h;
  Q = s0 + rem % h;
  return P < nsr;
}

__global__ void __launch_bounds__(NTHREADS, 1) k_trtri_pass1(SlotCtx c, int h) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  const int act = c.active[s], nb = c.nb[s];
  if (!act) return;
  const int nsr = (nb + 1) >> 1;
  int P, Q, mid;
  if (!trtri_decode(blockIdx.x, h, nsr, P, Q, mid)) return;
  double* Lt = c.Lt + (long)s * c.tile_stride;
  double* Xt = c.Xt + (long)s * c.tile_stride;
  Frag2H f(2 * P, nb - 1, (c.n[s] - (nb - 1) * TB) < 32);     // padded rows of the last tile row are not computed
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  gemm2_pipeline<false, true>(
      acc, smem, pipe, 2 * Q, 2 * mid,
      [&](int k, int t) -> const double* { return (2 * P + t < nb) ? tile_ptr(Lt, 2 * P + t, k) : nullptr; },
      [&](int k, int t) -> const double* { return (k >= 2 * Q + t) ? x_tile(c, s, k, 2 * Q + t) : nullptr; }, f);
  const int ti = 2 * P + f.ta, tj = 2 * Q + f.tb;
  if (ti < nb) store_acc2(tile_ptr(Xt, ti, tj), acc, f);
}

__global__ void __launch_bounds__(NTHREADS, 1) k_trtri_pass2(SlotCtx c, int h) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  const int act = c.active[s], nb = c.nb[s];
  if (!act) return;
  const int nsr = (nb + 1) >> 1;
  int P, Q, mid;
  if (!trtri_decode(blockIdx.x, h, nsr, P, Q, mid)) return;
  double* Lt = c.Lt + (long)s * c.tile_stride;
  double* Xt = c.Xt + (long)s * c.tile_stride;
  Frag2H f(2 * P, nb - 1, (c.n[s] - (nb - 1) * TB) < 32);
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  const int kend = (2 * P + 2 < nb) ? 2 * P + 2 : nb;
  // the A operand of the last k-tile(s) is the lower-TRIANGULAR diagonal tile X[2P+ta, 2P+ta]: its columns 32-63
  // (slice 1) are zero in rows 0-31
  const DiagRows rows{2 * P, 1, 2};
  gemm2_pipeline<false, true>(
      acc, smem, pipe, 2 * mid, kend,
      [&](int k, int t) -> const double* {
        return (2 * P + t < nb && k <= 2 * P + t) ? x_tile(c, s, 2 * P + t, k) : nullptr;
      },
      [&](int k, int t) -> const double* { return tile_ptr(Xt, k, 2 * Q + t); }, f, 0, rows);
  const int ti = 2 * P + f.ta, tj = 2 * Q + f.tb;
  if (ti < nb) store_acc2(tile_ptr(Lt, ti, tj), acc, f, -1.0);
}

// lauum: supertile (I, J), I >= J, of X'X -> Kt.   grid (nsr_max (nsr_max + 1) / 2, S)
__global__ void __launch_bounds__(NTHREADS, 1) k_lauum2(SlotCtx c) {
  extern __shared__ __align__(128) double smem[];
  const int s = blockIdx.y;
  const int act = c.active[s], nb = c.nb[s], N = c.n[s];
  if (!act) return;
  const int nsr = (nb + 1) >> 1;
  int I, J;
  tri_decode(blockIdx.x, I, J);
  if (I >= nsr) return;
  double* Kt = c.Kt + (long)s * c.tile_stride;
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  auto a_of = [&](int k, int t) -> const double* {
    return (2 * I + t < nb && k >= 2 * I + t) ? x_tile(c, s, k, 2 * I + t) : nullptr;
  };
  auto b_of = [&](int k, int t) -> const double* {
    return (2 * J + t < nb && k >= 2 * J + t) ? x_tile(c, s, k, 2 * J + t) : nullptr;
  };
  // rows N+1 .. 64 nb - 1 of X are identity padding: when they fill the second half of the last k-tile that slice adds
  // nothing to any entry the gradient reads, and it is not streamed (every task's k-loop ends with it)
  const int drop = ((N - (nb - 1) * TB) < 32) ? 1 : 0;
  // the A operand of the first k-tile(s) is the lower-TRIANGULAR diagonal tile X[2I+ta, 2I+ta] used transposed: its rows
  // 0-31 (slice 0) reach output rows 0-31 only
  const DiagRows rows{2 * I, 0, 1};
  if (I == J) {      // diagonal supertile: three tiles, balanced warp map (3/4 of a full supertile's time per k-step)
    Frag2D f;
    gemm2_pipeline<true, true>(acc, smem, pipe, 2 * I, nb, a_of, b_of, f, drop, rows);
    const int ti = 2 * I + f.ta, tj = 2 * J + f.tb;
    if (ti < nb && tj < nb) store_acc2(tile_ptr(Kt, ti, tj), acc, f);
    return;
  }
  Frag2 f;
  gemm2_pipeline<true, true>(acc, smem, pipe, 2 * I, nb, a_of, b_of, f, drop, rows);
  const int ti = 2 * I + f.ta, tj = 2 * J + f.tb;
  if (ti < nb && tj < nb && tj <= ti) store_acc2(tile_ptr(Kt, ti, tj), acc, f);
}

// gradient contraction per lower tile: reads (X'X)_ij from Kt, alpha from the augmented row of X.
// grid (ntmax, S), 256 threads: thread = (row, 16-column group)
template <int KID>
__global__ void __launch_bounds__(256, 4) k_grad_trace(SlotCtx c) {
  __shared__ double xi[MAXD * TB], xj[MAXD * TB], ai[TB], aj[TB], red[NG * 8];
  const int s = blockIdx.y;
  const int act = c.active[s], nb = c.nb[s], N = c.n[s];
  if (!act) return;
  int i, j;
  tri_decode(blockIdx.x, i, j);
  if (i >= nb) return;
  const double* th = c.theta + s * MAXP;
  const double* cs = c.coords + (long)s * MAXD * c.npmax;
  stage_coords(xi, cs, c.npmax, c.D, th, i, N);
  stage_coords(xj, cs, c.npmax, c.D, th, j, N);
  const int bN = nb - 1, rN = N - bN * TB;
  if (threadIdx.x < TB) {
    ai[threadIdx.x] = -x_tile(c, s, bN, i)[swz(rN, threadIdx.x)];
  } else if (threadIdx.x < 2 * TB) {
    const int m = threadIdx.x - TB;
    aj[m] = -x_tile(c, s, bN, j)[swz(rN, m)];
  }
  __syncthreads();
  const double kvar = th[c.D];
  const int m = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, gi = i * TB + m;
  const double* w = c.Kt + (long)s * c.tile_stride + tri_index(i, j) * TILE_ELEMS;
  double g[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) g[k] = 0.0;
  const bool interior = (i < nb - 1) && (j < i);   // all rows / columns are observations, nothing on the diagonal
  if (gi < N) {
    double xm[MAXD];
#pragma unroll
    for (int d = 0; d < MAXD; ++d) xm[d] = (d < c.D) ? xi[d * TB + m] : 0.0;
    const double am2 = 2.0 * ai[m];
#pragma unroll 1
    for (int cc = 0; cc < 16; cc += 4) {
      double wv[4];
      ld_global_v4(w + swz(m, c0 + cc), wv[0], wv[1], wv[2], wv[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = c0 + cc + e, gj = j * TB + n;
        if (interior || gj < N) {
          const double W = wv[e] - am2 * aj[n];
          double r2 = 0.0, d2[MAXD];
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            d2[d] = 0.0;
            if (d < c.D) {
              const double df = xm[d] - xj[d * TB + n];
              d2[d] = df * df;
              r2 += d2[d];
            }
          }
          double kv, hv;
          kern_eval_t<KID>(r2, kvar, kv, hv);
          const double wh = W * hv;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) g[d] += wh * d2[d];
          g[MAXD] += W * kv;
          if (!interior && gi == gj) g[MAXD + 1] += W;
        }
      }
    }
  }
  const double wgt = (i != j) ? 2.0 : 1.0;
  block_sum<NG>(g, red);
  if (threadIdx.x == 0) {
    double* gp = c.gpart + ((long)s * c.ntmax + blockIdx.x) * NG;
#pragma unroll
    for (int k = 0; k < NG; ++k) gp[k] = wgt * g[k];
  }
}

// finalize: -LML and d(-LML)/dtheta per slot.  grid (S), NTHREADS threads
__global__ void __launch_bounds__(NTHREADS) k_finalize2(SlotCtx c, int with_grad) {
  __shared__ double red[(NG + 1) * (NTHREADS / 32)];
  const int s = blockIdx.x;
  if (!c.active[s]) return;
  const int N = c.n[s], nb = c.nb[s];
  double v[NG + 1];
#pragma unroll
  for (int k = 0; k < NG + 1; ++k) v[k] = 0.0;
  for (int k = threadIdx.x; k < nb; k += NTHREADS) v[NG] += c.logdet_part[s * c.nbmax + k];
  if (with_grad) {
    const int nt = nb * (nb + 1) / 2;
    for (int t = threadIdx.x; t < nt; t += NTHREADS) {
      const double* gp = c.gpart + ((long)s * c.ntmax + t) * NG;
#pragma unroll
      for (int k = 0; k < NG; ++k) v[k] += gp[k];
    }
  }
  block_sum<NG + 1>(v, red);
  if (threadIdx.x == 0) {
    const double* th = c.theta + s * MAXP;
    double f = 0.5 * c.quad[s] + v[NG] + 0.5 * N * 1.8378770664093453;
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
// Prediction.  Work item = (slot, pair of 64-wide blocks of prediction points).
//   k_build_xp : cross-covariance tiles K(x_obs block k, x_pred block) -> scratch[item][k][t]
//   k_predict2 : one CTA per item walks the row supertiles: A_I = sum_k X_{I,k} Kxp_k;
//                column sums of A^2 over rows < N; row N -> -mean
// ------------------------------------------------------------------------------------
struct PredCtx {
  int ppmax;                 // padded prediction points per slot
  const double* pcoords;     // [S][MAXD][ppmax] prediction coords / coords_scale
  const int* np;             // [S]
  const int* item_slot;      // [n_items]
  const int* item_pb;        // [n_items]  first 64-block of the pair
  int n_items, item0;        // items of this wave: [item0, item0 + n_items)
  double* scratch;           // [n_items][nbmax][2][4096]
  double* fmean;             // [S][ppmax]
  double* fvar;              // [S][ppmax]
  double* abuf;              // [n_items][nbmax][2][4096] A = L^-1 K_xp tiles (full_cov only) or nullptr
};

// grid (nbmax, n_items), 256 threads
__global__ void __launch_bounds__(256, 4) k_build_xp(SlotCtx c, PredCtx p) {
  __shared__ double xo[MAXD * TB], xp[MAXD * 2 * TB];
  const int item = p.item0 + blockIdx.y, k = blockIdx.x;
  const int s = p.item_slot[item], pb = p.item_pb[item];
  const int nb = c.nb[s];
  if (k >= nb) return;
  const int N = c.n[s], P = p.np[s];
  const double* th = c.theta + s * MAXP;
  stage_coords(xo, c.coords + (long)s * MAXD * c.npmax, c.npmax, c.D, th, k, N);
  const double* ps = p.pcoords + (long)s * MAXD * p.ppmax;
  for (int t = threadIdx.x; t < c.D * 2 * TB; t += blockDim.x) {
    const int d = t / (2 * TB), m = t % (2 * TB), g = pb * TB + m;
    xp[d * 2 * TB + m] = (g < P) ? ps[(long)d * p.ppmax + g] / th[d] : 0.0;
  }
  __syncthreads();
  const double kvar = th[c.D];
  const int kk = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, go = k * TB + kk;
  double* out = p.scratch + ((long)blockIdx.y * c.nbmax + k) * 2 * TILE_ELEMS;
  double xm[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) xm[d] = (d < c.D) ? xo[d * TB + kk] : 0.0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int cc = 0; cc < 16; cc += 2) {
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = t * TB + c0 + cc + e;
        double val = 0.0;
        if (go < N && pb * TB + n < P) {
          double r2 = 0.0;
#pragma unroll
          for (int d = 0; d < MAXD; ++d)
            if (d < c.D) {
              const double df = xm[d] - xp[d * 2 * TB + n];
              r2 += df * df;
            }
          val = kern_value(c.kid, r2, kvar);
        }
        v[e] = val;
      }
      *reinterpret_cast<double2*>(out + t * TILE_ELEMS + swz(kk, c0 + cc)) = make_double2(v[0], v[1]);
    }
  }
}

// grid (n_items), NTHREADS threads
template <bool STORE_A>
__global__ void __launch_bounds__(NTHREADS, 1) k_predict2(SlotCtx c, PredCtx p) {
  extern __shared__ __align__(128) double smem[];
  double* meanv = smem + (SMEM2_ELEMS - G2_AUX);      // [128]
  const int item = p.item0 + blockIdx.x;
  const int s = p.item_slot[item], pb = p.item_pb[item];
  const int N = c.n[s], nb = c.nb[s], nsr = (nb + 1) >> 1, P = p.np[s];
  const bool two = (pb + 1) * TB < P;
  const double* scr = p.scratch + (long)blockIdx.x * c.nbmax * 2 * TILE_ELEMS;
  const double kvar = c.theta[s * MAXP + c.D];
  Frag2 f;
  G2Pipe pipe;
  pipe.init();
  double csq[4][2];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) csq[ni][0] = csq[ni][1] = 0.0;
  const int bN = nb - 1, rN = N - bN * TB;
  for (int I = 0; I < nsr; ++I) {
    Acc2 acc;
    acc.zero();
    const int kend = (2 * I + 2 < nb) ? 2 * I + 2 : nb;
    gemm2_pipeline<false, true>(
        acc, smem, pipe, 0, kend,
        [&](int k, int t) -> const double* {
          return (2 * I + t < nb && k <= 2 * I + t) ? x_tile(c, s, 2 * I + t, k) : nullptr;
        },
        [&](int k, int t) -> const double* {
          return (t == 0 || two) ? scr + ((long)k * 2 + t) * TILE_ELEMS : nullptr;
        },
        f, 0, DiagRows{2 * I, 1, 2});     // columns 32-63 of the triangular tile X[2I+t, 2I+t] reach rows 32-63 only
    const int ti = 2 * I + f.ta;
    if (ti < nb && (f.tb == 0 || two)) {
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int gi = ti * TB + f.row(mi);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if (gi < N) {
            csq[ni][0] += acc.c[mi][ni][0] * acc.c[mi][ni][0];
            csq[ni][1] += acc.c[mi][ni][1] * acc.c[mi][ni][1];
          } else if (ti == bN && f.row(mi) == rN) {
            meanv[f.tb * TB + f.col(ni)] = -acc.c[mi][ni][0];
            meanv[f.tb * TB + f.col(ni) + 1] = -acc.c[mi][ni][1];
          }
        }
      }
    }
    if (STORE_A && ti < nb) {   // rows >= N (augmented row, padding) must not enter A'A: zeroed in place, then stored
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
        if (ti * TB + f.row(mi) >= N || !(f.tb == 0 || two)) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) acc.c[mi][ni][0] = acc.c[mi][ni][1] = 0.0;
        }
      store_acc2(p.abuf + (((long)blockIdx.x * c.nbmax + ti) * 2 + f.tb) * TILE_ELEMS, acc, f);
    }
  }
  // column sums: over q (lanes sharing r), then over the two warps (ta = 0, 1) of a column slab
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
  double* part = smem;   // [2][128]: the pipeline ended with a barrier, the ring is idle
  if (f.q == 0) {
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      part[f.ta * 2 * TB + f.tb * TB + f.col(ni)] = csq[ni][0];
      part[f.ta * 2 * TB + f.tb * TB + f.col(ni) + 1] = csq[ni][1];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * TB) {
    const int n = threadIdx.x, gp = pb * TB + n;
    if (gp < P) {
      p.fvar[(long)s * p.ppmax + gp] = kvar - (part[n] + part[2 * TB + n]);
      p.fmean[(long)s * p.ppmax + gp] = meanv[n];
    }
  }
}

// full posterior covariance of one expert (slot 0): C(p, q) = K(x*_p, x*_q) - sum_k A_k,p' A_k,q.
// grid (n_items (n_items + 1) / 2): lower 128x128 blocks, mirrored on store.  fcov: [P][P] row-major
__global__ void __launch_bounds__(NTHREADS, 1) k_pred_cov(SlotCtx c, PredCtx p, double* __restrict__ fcov) {
  extern __shared__ __align__(128) double smem[];
  int bp, bq;
  tri_decode(blockIdx.x, bp, bq);
  const int s = 0, nb = c.nb[s], P = p.np[s];
  Frag2 f;
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  const long istride = (long)c.nbmax * 2 * TILE_ELEMS;
  gemm2_pipeline<true, true>(
      acc, smem, pipe, 0, nb,
      [&](int k, int t) -> const double* { return p.abuf + bp * istride + ((long)k * 2 + t) * TILE_ELEMS; },
      [&](int k, int t) -> const double* { return p.abuf + bq * istride + ((long)k * 2 + t) * TILE_ELEMS; }, f);
  const double* th = c.theta + s * MAXP;
  const double* ps = p.pcoords + (long)s * MAXD * p.ppmax;
  const double kvar = th[c.D];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int gp = (2 * bp + f.ta) * TB + f.row(mi);
    if (gp >= P) continue;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gq = (2 * bq + f.tb) * TB + f.col(ni) + e;
        if (gq >= P) continue;
        double r2 = 0.0;
        for (int d = 0; d < c.D; ++d) {
          const double df = ps[(long)d * p.ppmax + gp] / th[d] - ps[(long)d * p.ppmax + gq] / th[d];
          r2 += df * df;
        }
        const double v = kern_value(c.kid, r2, kvar) - acc.c[mi][ni][e];
        fcov[(long)gp * P + gq] = v;
        fcov[(long)gq * P + gp] = v;
      }
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

// unpack packed swizzled lower tiles of one slot into a dense row-major (npad x npad) matrix.
// which = 0: Lt as is; 1: X assembled through x_tile()
__global__ void k_unpack_tiles(SlotCtx c, int s, int which, int nb, double* __restrict__ dense) {
  const int i = blockIdx.y, j = blockIdx.x;
  const int npad = nb * TB;
  const double* src = nullptr;
  if (j <= i) src = which ? x_tile(c, s, i, j) : tile_ptr(c.Lt + (long)s * c.tile_stride, i, j);
  for (int t = threadIdx.x; t < TILE_ELEMS; t += blockDim.x) {
    const int r = t >> 6, cc = t & 63;
    dense[(long)(i * TB + r) * npad + j * TB + cc] = src ? src[swz(r, cc)] : 0.0;
  }
}

}  // namespace gpsat
