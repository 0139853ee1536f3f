// gpsat_b200: common device primitives (sm_100a).
//
// Tile format used by every dense kernel in this library
// ------------------------------------------------------
// A symmetric / triangular N x N matrix of one expert is stored as PACKED LOWER-TRIANGULAR
// 64 x 64 TILES: tile (i, j), j <= i, lives at tile index i*(i+1)/2 + j and is one contiguous
// 32 KiB blob made of two 64 x 32 COLUMN HALVES (16 KiB each, columns 0-31 then 32-63).
// Inside a tile element (r, c) is stored at
//       (c >> 5) * 2048 + r * 32 + ((c & 31) ^ ((r & 3) << 2))        (doubles).
// Consequences:
//   * a 32-deep k-slice of an operand is contiguous in global memory whichever way k runs: along the
//     columns it is one column half (one 16 KiB TMA bulk copy), along the rows it is rows
//     [32h, 32h+32) of both halves (two 8 KiB bulk copies) -- no per-thread cp.async traffic at all;
//   * the XOR swizzle makes BOTH DMMA fragment access patterns bank-conflict free straight from
//     the copied image (no padding, no re-layout pass):
//       row pattern  : lane reads (row0 + lane/4, k0 + lane%4)
//       col pattern  : lane reads (k0 + lane%4, col0 + lane/4)
//     (a 64-bit shared load is served per half-warp over 16 8-byte banks; in both patterns the
//      16 lanes of a half-warp hit 16 distinct banks.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpsat {

constexpr int TB = 64;                 // tile edge
constexpr int TILE_ELEMS = TB * TB;    // 4096 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int MAXD = 4;                // max coordinate dimension
constexpr int MAXP = MAXD + 2;         // lengthscales..., kernel variance, likelihood variance
constexpr int NTHREADS = 256;          // CTA size of the tile kernels (8 warps, two per scheduler)

enum KernelId { K_MATERN32 = 0, K_MATERN52 = 1, K_MATERN12 = 2, K_RBF = 3 };

constexpr int HALF_ELEMS = TB * 32;    // one 64 x 32 column half = 2048 doubles = 16 KiB
__host__ __device__ __forceinline__ int swz(int r, int c) {
  return ((c >> 5) << 11) + r * 32 + ((c & 31) ^ ((r & 3) << 2));
}
__host__ __device__ __forceinline__ long tri_index(int i, int j) { return (long)i * (i + 1) / 2 + j; }

// ---- cp.async (LDGSTS) 16-byte copies (micro-benchmark baseline only; the kernels use the TMA helpers below) ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// copy one 32 KiB tile global -> shared with all NTHREADS threads (8 x 16 B each)
__device__ __forceinline__ void load_tile_async(double* smem_tile, const double* gmem_tile) {
#pragma unroll
  for (int c = 0; c < TILE_BYTES / 16 / NTHREADS; ++c) {
    int idx = threadIdx.x + c * NTHREADS;
    cp_async16(reinterpret_cast<char*>(smem_tile) + idx * 16,
               reinterpret_cast<const char*>(gmem_tile) + idx * 16);
  }
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP; completion counted on an mbarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// orders this thread's generic-proxy accesses (e.g. an acquire load of a flag another CTA released after writing
// tiles to global memory) before its subsequent async-proxy (TMA) reads of that global data
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one thread: copy `bytes` (multiple of 16, both addresses 16-byte aligned) global -> shared
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per lane and instruction ----
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_global_v4(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8) ----
// A: lane holds A[lane/4][lane%4]; B: lane holds B[lane%4][lane/4]; C: lane holds C[lane/4][2*(lane%4)+{0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- warp / block reductions ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum of NV values per thread; result valid in thread 0. red: >= NV*8 doubles
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) red[k * (NTHREADS / 32) + warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < NTHREADS / 32; ++w) s += red[k * (NTHREADS / 32) + w];
      v[k] = s;
    }
  }
  __syncthreads();
}

// ---- lean elementary functions for the kernel-matrix kernels ----
// The DFMA and DMMA pipes are the same hardware on sm_100, so every FP64 instruction of the elementwise kernels is
// paid in tensor time.  libdevice's exp() / sqrt() carry range checks and slow paths these call sites never need
// (the argument of the exponential is always <= 0, r2 is clamped to >= 1e-36): ncu counted 111 instructions per
// matrix entry with them.
// exp(-u), u >= 0: n = rint(-u log2 e), f = -u - n ln2 (two-term Cody-Waite), degree-13 Taylor on |f| <= 0.347
// (truncation 4e-18), scaled by 2^n through the exponent field; 0 beyond u = 700 (true value < 1e-304).
__device__ __forceinline__ double exp_neg(double u) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the low word of (x + MAGIC) is rint(x)
  const double t = fma(-u, 1.4426950408889634, MAGIC);
  const int n = __double2loint(t);
  const double nd = t - MAGIC;
  double f = fma(nd, -6.93147180369123816490e-01, -u);
  f = fma(nd, -1.90821492927058770002e-10, f);
  double p = 1.6059043836821613e-10;          // 1/13!
  p = fma(p, f, 2.08767569878681e-09);        // 1/12!
  p = fma(p, f, 2.505210838544172e-08);       // 1/11!
  p = fma(p, f, 2.755731922398589e-07);       // 1/10!
  p = fma(p, f, 2.7557319223985893e-06);      // 1/9!
  p = fma(p, f, 2.48015873015873e-05);        // 1/8!
  p = fma(p, f, 1.984126984126984e-04);       // 1/7!
  p = fma(p, f, 1.388888888888889e-03);       // 1/6!
  p = fma(p, f, 8.333333333333333e-03);       // 1/5!
  p = fma(p, f, 4.1666666666666664e-02);      // 1/4!
  p = fma(p, f, 1.6666666666666666e-01);      // 1/3!
  p = fma(p, f, 0.5);
  p = fma(p, f, 1.0);
  p = fma(p, f, 1.0);
  const double r = __hiloint2double(__double2hiint(p) + n * 1048576, __double2loint(p));
  return (u > 700.0) ? 0.0 : r;
}
// sqrt(x) for x >= 1e-36 (normal range, no special cases): x * rsqrt(x)
__device__ __forceinline__ double sqrt_pos(double x) { return x * rsqrt(x); }

// ---- stationary kernel functions (gpflow.kernels.*) ----
// k(r2) and h(r2) with dk/dl_d = h * delta_d^2 / l_d^3   (delta in unscaled units)
template <int KID>
__device__ __forceinline__ void kern_eval_t(double r2, double var, double& k, double& h) {
  if (KID == K_RBF) {
    k = var * exp_neg(0.5 * r2);
    h = k;
    return;
  }
  const double r2c = fmax(r2, 1e-36);
  const double r = sqrt_pos(r2c);
  if (KID == K_MATERN32) {
    const double s3 = 1.7320508075688772;
    const double e = exp_neg(s3 * r);
    k = var * (1.0 + s3 * r) * e;
    h = 3.0 * var * e;
  } else if (KID == K_MATERN52) {
    const double s5 = 2.23606797749979;
    const double e = exp_neg(s5 * r);
    k = var * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * e;
    h = var * (5.0 / 3.0) * (1.0 + s5 * r) * e;
  } else {  // Matern12 / Exponential
    const double e = exp_neg(r);
    k = var * e;
    h = (r2 > 1e-36) ? k / r : 0.0;
  }
}
__device__ __forceinline__ void kern_eval(int kid, double r2, double var, double& k, double& h) {
  switch (kid) {
    case K_RBF: kern_eval_t<K_RBF>(r2, var, k, h); break;
    case K_MATERN32: kern_eval_t<K_MATERN32>(r2, var, k, h); break;
    case K_MATERN52: kern_eval_t<K_MATERN52>(r2, var, k, h); break;
    default: kern_eval_t<K_MATERN12>(r2, var, k, h); break;
  }
}
template <int KID>
__device__ __forceinline__ double kern_value_t(double r2, double var) {
  double k, h;
  kern_eval_t<KID>(r2, var, k, h);
  return k;
}
__device__ __forceinline__ double kern_value(int kid, double r2, double var) {
  double k, h;
  kern_eval(kid, r2, var, k, h);
  return k;
}

}  // namespace gpsat
