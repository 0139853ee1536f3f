// gpsat_b200: 128x64 FP64 GEMM core for TWO co-resident CTAs per SM (experimental, microbenchmark only so far).
//
// Same DMMA fragment code as gemm2.cuh (every warp owns a 64x32 accumulator slab) but a CTA has 4 warps -- one per
// scheduler -- and owns a 2x1 group of 64x64 output tiles; the ring has 2 slots of 48 KiB (A0, A1, B halves of a
// 32-deep k-slice), i.e. 96 KiB per CTA, and 254 registers x 128 threads x 2 CTAs fill the register file exactly.
// With two CTAs on an SM the prologue / epilogue / diagonal-block work of one overlaps the k-loop of the other:
// the measurement (profiles/r01_microbench_2cta.json) decides whether the hot kernels move to this shape.
#pragma once
#include "gemm2.cuh"

namespace gpsat {

constexpr int G3_THREADS = 128;
constexpr int G3_STAGES = 2;
constexpr int G3_STAGE_ELEMS = 3 * HALF_ELEMS;                 // A0 A1 B
constexpr int G3_SMEM_ELEMS = G3_STAGES * G3_STAGE_ELEMS;      // 12288 doubles = 96 KiB

struct Frag3 {
  int lane, warp, q, r, ta, nb0;
  __device__ __forceinline__ Frag3() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    q = lane >> 2;
    r = lane & 3;
    ta = warp >> 1;
    nb0 = (warp & 1) * 32;
  }
  __device__ __forceinline__ int row(int mi) const { return 8 * mi + q; }
  __device__ __forceinline__ int col(int ni) const { return nb0 + 8 * ni + 2 * r; }
};

struct G3Pipe {
  uint64_t* full;
  int* done;
  uint32_t count;
  __device__ __forceinline__ void init() {
    __shared__ __align__(8) uint64_t bars[G3_STAGES];
    __shared__ int cnt[G3_STAGES];
    full = bars;
    done = cnt;
    count = 0;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < G3_STAGES; ++s) {
        mbar_init(bars + s, 1);
        cnt[s] = 0;
      }
      fence_mbar_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ int stage(uint32_t n) const { return (int)(n % G3_STAGES); }
  __device__ __forceinline__ void wait(uint32_t n) const { mbar_wait(full + stage(n), (n / G3_STAGES) & 1u); }
};

// acc += sum_k [A_k^0; A_k^1] * B_k ; a_of(k, t), t in {0, 1}, b_of(k): tile pointers or nullptr (CTA-uniform)
template <bool TA, bool TBm, class FA, class FB>
__device__ __forceinline__ void gemm3_pipeline(Acc2& acc, double* smem, G3Pipe& p, int kbeg, int kend, FA a_of, FB b_of,
                                               const Frag3& f) {
  const int nsl = (kend > kbeg) ? 2 * (kend - kbeg) : 0;
  if (nsl == 0) return;
  auto issue = [&](int sl) {
    const uint32_t n = p.count + sl;
    double* st = smem + p.stage(n) * G3_STAGE_ELEMS;
    uint64_t* bar = p.full + p.stage(n);
    const int k = kbeg + (sl >> 1), kh = sl & 1;
    const double* pa0 = a_of(k, 0);
    const double* pa1 = a_of(k, 1);
    const double* pb = b_of(k);
    mbar_expect_tx(bar, 16384u * ((pa0 != nullptr) + (pa1 != nullptr) + (pb != nullptr)));
    if (pa0) bulk_half<TA>(st, pa0, kh, bar);
    if (pa1) bulk_half<TA>(st + HALF_ELEMS, pa1, kh, bar);
    if (pb) bulk_half<TBm>(st + 2 * HALF_ELEMS, pb, kh, bar);
  };
  if (threadIdx.x == 0) {
    for (int sl = 0; sl < G3_STAGES && sl < nsl; ++sl) issue(sl);
  }
  for (int sl = 0; sl < nsl; ++sl) {
    const uint32_t n = p.count + sl;
    p.wait(n);
    const int k = kbeg + (sl >> 1);
    if (a_of(k, f.ta) != nullptr && b_of(k) != nullptr) {
      const double* st = smem + p.stage(n) * G3_STAGE_ELEMS;
      mma_half<TA, TBm>(acc, st + f.ta * HALF_ELEMS, st + 2 * HALF_ELEMS, f);
    }
    if (sl + G3_STAGES < nsl) {
      __syncwarp();
      if (f.lane == 0) {
        int* cnt = p.done + p.stage(n);
        if (atomicAdd(cnt, 1) == G3_THREADS / 32 - 1) {
          atomicExch(cnt, 0);
          __threadfence_block();
          issue(sl + G3_STAGES);
        }
      }
    }
  }
  __syncthreads();
  p.count += nsl;
}

}  // namespace gpsat
