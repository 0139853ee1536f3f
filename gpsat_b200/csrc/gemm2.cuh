// gpsat_b200: 128x128 "supertile" FP64 GEMM core on the DMMA pipe (mma.sync m8n8k4 f64).
//
// One CTA (8 warps, 2 along M x 4 along N, 64x32 per warp) owns a 2x2 group of 64x64 output
// tiles and streams operand tiles (packed swizzled 32 KiB blobs, see common.cuh) through a
// 3-stage cp.async ring of 32-deep k-slices:  per slice 4 half-tiles (2 of A, 2 of B) = 64 KiB,
// 128x128x32 FMAs -> 32 KiB of operand traffic per 64^3 tile product (half of the 64x64 core).
// Fragments are double-buffered in registers so the shared-memory loads of step k+4 are in
// flight while the 32 DMMAs of step k issue.
//
// Operand orientation (per tile, straight from the swizzled image):
//   TA  = false: A[m][k] = Atile(m, k)  (k along tile columns)   TA  = true: A[m][k] = Atile(k, m)
//   TBm = false: B[k][n] = Btile(n, k)                            TBm = true: B[k][n] = Btile(k, n)
// A k-slice of a "k along columns" tile is the 64 x 32 column half (row stride 32 in shared
// memory, the XOR swizzle only touches column bits 2-3 so it survives the split); a k-slice of a
// "k along rows" tile is 32 full rows (contiguous 16 KiB).
#pragma once
#include "common.cuh"

namespace gpsat {

constexpr int G2_STAGES = 3;
constexpr int HALF_ELEMS = TB * 32;                  // 2048 doubles = 16 KiB
constexpr int G2_STAGE_ELEMS = 4 * HALF_ELEMS;       // A0 A1 B0 B1
constexpr int G2_SMEM_ELEMS = G2_STAGES * G2_STAGE_ELEMS;   // 24576 doubles = 192 KiB

struct Acc2 {
  double c[8][4][2];   // [mi][ni][pair]: row 8*mi + q of A-tile wm; col (wn&1)*32 + 8*ni + 2*r + {0,1} of B-tile wn>>1
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  }
};

struct Frag2 {
  int lane, warp, q, r, wm, wn, ta, tb, nb0;
  __device__ __forceinline__ Frag2() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    q = lane >> 2;
    r = lane & 3;
    wm = warp & 1;        // which A tile (row tile of the 2x2 group)
    wn = warp >> 1;       // 0..3: 32-column slab
    ta = wm;
    tb = wn >> 1;         // which B tile (column tile of the 2x2 group)
    nb0 = (wn & 1) * 32;  // column offset inside that tile
  }
  __device__ __forceinline__ int row(int mi) const { return 8 * mi + q; }             // within tile ta
  __device__ __forceinline__ int col(int ni) const { return nb0 + 8 * ni + 2 * r; }   // within tile tb (and col+1)
};

// issue the cp.async copies of one half-tile (16 KiB) with all NTHREADS threads
template <bool KROWS>   // KROWS: k runs along tile rows (contiguous slice); else along columns
__device__ __forceinline__ void load_half_async(double* smem_half, const double* gmem_tile, int kh) {
  const char* src = reinterpret_cast<const char*>(gmem_tile);
  char* dst = reinterpret_cast<char*>(smem_half);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int idx = threadIdx.x + c * NTHREADS;   // 0..1023 chunks of 16 B
    if (KROWS) {
      cp_async16(dst + idx * 16, src + kh * 16384 + idx * 16);
    } else {
      const int rr = idx >> 4, x = idx & 15;
      cp_async16(dst + rr * 256 + x * 16, src + rr * 512 + kh * 256 + x * 16);
    }
  }
}

// acc += op(A_half) * op(B_half) for this warp's 64x32 slab, 32-deep slice resident in shared memory
template <bool TA, bool TBm>
__device__ __forceinline__ void mma_half(Acc2& acc, const double* __restrict__ As, const double* __restrict__ Bs,
                                         const Frag2& f) {
  int aoff[8], boff[4];
  const int sq = (f.q & 3) << 2;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int m = 8 * mi + f.q;
    aoff[mi] = TA ? (f.r * TB + (m ^ (f.r << 2))) : (m * 32 + f.r);
  }
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    const int n = f.nb0 + 8 * ni + f.q;
    boff[ni] = TBm ? (f.r * TB + (n ^ (f.r << 2))) : (n * 32 + f.r);
  }
  double a[2][8], b[2][4];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) a[0][mi] = As[aoff[mi] + (TA ? 0 : (0 ^ sq))];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) b[0][ni] = Bs[boff[ni] + (TBm ? 0 : (0 ^ sq))];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int cur = ks & 1, nxt = cur ^ 1;
    if (ks + 1 < 8) {
      const int kk = (ks + 1) * 4;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) a[nxt][mi] = As[aoff[mi] + (TA ? kk * TB : (kk ^ sq))];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[nxt][ni] = Bs[boff[ni] + (TBm ? kk * TB : (kk ^ sq))];
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc.c[mi][ni][0], acc.c[mi][ni][1], a[cur][mi], b[cur][ni]);
  }
}

// acc += sum_{k = kbeg}^{kend-1} [A_k^0; A_k^1] * [B_k^0  B_k^1]  over 64-deep tile steps.
// a_of(k, t) / b_of(k, t), t in {0,1}: global pointer of the tile or nullptr (structurally zero /
// out of range: the copy and the products that would use it are skipped; nullness must be
// CTA-uniform).  smem: G2_SMEM_ELEMS doubles.  Ends with all copies drained and a __syncthreads().
template <bool TA, bool TBm, class FA, class FB>
__device__ __forceinline__ void gemm2_pipeline(Acc2& acc, double* smem, int kbeg, int kend, FA a_of, FB b_of,
                                               const Frag2& f) {
  const int nsl = 2 * (kend - kbeg);   // 32-deep slices
  if (nsl <= 0) return;
  auto issue = [&](int sl) {
    const int k = kbeg + (sl >> 1), kh = sl & 1;
    double* st = smem + (sl % G2_STAGES) * G2_STAGE_ELEMS;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const double* pa = a_of(k, t);
      if (pa) load_half_async<TA>(st + t * HALF_ELEMS, pa, kh);
      const double* pb = b_of(k, t);
      if (pb) load_half_async<TBm>(st + (2 + t) * HALF_ELEMS, pb, kh);
    }
  };
  issue(0);
  cp_async_commit();
  if (nsl > 1) issue(1);
  cp_async_commit();
  for (int sl = 0; sl < nsl; ++sl) {
    cp_async_wait<1>();      // slice sl has landed (one younger group may be in flight)
    __syncthreads();         // ... for every thread; and everyone is done with slice sl-1's buffer
    if (sl + 2 < nsl) issue(sl + 2);
    cp_async_commit();
    const int k = kbeg + (sl >> 1);
    if (a_of(k, f.ta) != nullptr && b_of(k, f.tb) != nullptr) {
      const double* st = smem + (sl % G2_STAGES) * G2_STAGE_ELEMS;
      mma_half<TA, TBm>(acc, st + f.ta * HALF_ELEMS, st + (2 + f.tb) * HALF_ELEMS, f);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// store this warp's 64x32 slab into a swizzled 64x64 tile (global or shared), scaled
__device__ __forceinline__ void store_acc2(double* __restrict__ tile, const Acc2& acc, const Frag2& f,
                                           double scale = 1.0) {
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const int m = f.row(mi);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int n = f.col(ni);
      *reinterpret_cast<double2*>(tile + swz(m, n)) =
          make_double2(scale * acc.c[mi][ni][0], scale * acc.c[mi][ni][1]);
    }
  }
}

}  // namespace gpsat
