// gpsat_b200: 64x64 FP64 tile GEMM core on the DMMA pipe (mma.sync m8n8k4 f64).
//
// tcgen05 has no f64 kind, so the FP64 tensor path on sm_100a is warp-level DMMA with
// register accumulators.  One CTA (8 warps, 4 along M x 2 along N, 16x32 per warp) owns one
// 64x64 output tile.  mma_tile() multiplies two tiles resident in shared memory; the production kernels use it
// for the 64^3 products inside the diagonal 128x128 blocks (gpr2.cuh: diag_block_128).  gemm_pipeline() -- the
// original 2-stage per-thread cp.async ring over whole tiles -- is kept only as the baseline of the
// micro-benchmarks ("core64" lines); everything hot streams through the TMA ring of gemm2.cuh.
// Both operands can be consumed in either orientation straight from the swizzled image:
//   TA = false : A[m][k] = Atile(m, k)        TA = true : A[m][k] = Atile(k, m)
//   TBm = false: B[k][n] = Btile(n, k)  (NT)  TBm = true: B[k][n] = Btile(k, n)
#pragma once
#include "common.cuh"

namespace gpsat {

struct Acc {
  double c[2][4][2];  // [mi][ni][pair]  rows 16*wm + 8*mi + q, cols 32*wn + 8*ni + 2*r + {0,1}
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  }
};

struct FragCoord {
  int lane, warp, q, r, wm, wn;
  __device__ __forceinline__ FragCoord() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    q = lane >> 2;
    r = lane & 3;
    wm = warp & 3;
    wn = warp >> 2;
  }
  __device__ __forceinline__ int row(int mi) const { return 16 * wm + 8 * mi + q; }
  __device__ __forceinline__ int col(int ni) const { return 32 * wn + 8 * ni + 2 * r; }  // and col+1
};

// acc += op(As) * op(Bs) for one pair of 64x64 tiles resident in shared memory
template <bool TA, bool TBm>
__device__ __forceinline__ void mma_tile(Acc& acc, const double* __restrict__ As,
                                         const double* __restrict__ Bs, const FragCoord& fc) {
  // tile image (common.cuh): (r, c) at (c >> 5) * 2048 + r * 32 + ((c & 31) ^ ((r & 3) << 2))
  int abase[2], bbase[4];
  const int sq = (fc.q & 3) << 2;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int m = 16 * fc.wm + 8 * mi + fc.q;
    abase[mi] = TA ? ((m >> 5) * HALF_ELEMS + fc.r * 32 + ((m & 31) ^ (fc.r << 2))) : (m * 32 + fc.r);
  }
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    const int n = 32 * fc.wn + 8 * ni + fc.q;
    bbase[ni] = TBm ? ((n >> 5) * HALF_ELEMS + fc.r * 32 + ((n & 31) ^ (fc.r << 2))) : (n * 32 + fc.r);
  }
#pragma unroll
  for (int k0 = 0; k0 < TB; k0 += 4) {
    double a[2], b[4];
    // k along columns: half k0 >> 5, column (k0 & 31) + r, swizzled by the row: ((k0 & 31) ^ sq) + r
    const int kx = (k0 >> 5) * HALF_ELEMS + ((k0 & 31) ^ sq);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) a[mi] = TA ? As[abase[mi] + k0 * 32] : As[abase[mi] + kx];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = TBm ? Bs[bbase[ni] + k0 * 32] : Bs[bbase[ni] + kx];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc.c[mi][ni][0], acc.c[mi][ni][1], a[mi], b[ni]);
  }
}

// acc += sum_{k = kbeg}^{kend-1} op(A_k) * op(B_k); a_of(k) / b_of(k) give global tile pointers.
// smem: 4 tiles (2 stages x {A, B}).  Ends with all async copies drained and a __syncthreads().
template <bool TA, bool TBm, class FA, class FB>
__device__ __forceinline__ void gemm_pipeline(Acc& acc, double* smem, int kbeg, int kend, FA a_of,
                                              FB b_of, const FragCoord& fc) {
  if (kbeg >= kend) return;
  load_tile_async(smem, a_of(kbeg));
  load_tile_async(smem + TILE_ELEMS, b_of(kbeg));
  cp_async_commit();
  for (int k = kbeg; k < kend; ++k) {
    const int st = (k - kbeg) & 1;
    if (k + 1 < kend) {
      double* nx = smem + (st ^ 1) * 2 * TILE_ELEMS;
      load_tile_async(nx, a_of(k + 1));
      load_tile_async(nx + TILE_ELEMS, b_of(k + 1));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    mma_tile<TA, TBm>(acc, smem + st * 2 * TILE_ELEMS, smem + st * 2 * TILE_ELEMS + TILE_ELEMS, fc);
    __syncthreads();
  }
}

// store the accumulator tile to a swizzled 64x64 tile (global or shared), scaled by `scale`
__device__ __forceinline__ void store_acc_swizzled(double* __restrict__ tile, const Acc& acc,
                                                   const FragCoord& fc, double scale = 1.0) {
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int m = fc.row(mi);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int n = fc.col(ni);
      double2 v = make_double2(scale * acc.c[mi][ni][0], scale * acc.c[mi][ni][1]);
      *reinterpret_cast<double2*>(tile + swz(m, n)) = v;
    }
  }
}

}  // namespace gpsat
