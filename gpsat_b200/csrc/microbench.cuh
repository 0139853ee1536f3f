// gpsat_b200: micro-benchmarks that establish the roofline denominators and the efficiency of the
// GEMM cores in isolation (used by bench.py / profiles, not by the product path).
#pragma once
#include "gemm_core.cuh"
#include "gemm2.cuh"
#include "gpr2.cuh"

namespace gpsat {

// register-resident DMMA chains: NCH independent accumulators per warp
template <int NCH>
__global__ void __launch_bounds__(256) k_dmma_chain(int iters, double* sink) {
  double c[NCH][2];
#pragma unroll
  for (int k = 0; k < NCH; ++k) c[k][0] = c[k][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) sink[0] = s;
}

// do the FP64 FMA pipe and the DMMA pipe run concurrently?  warps 0-3 of each CTA issue DMMA chains
// (mode & 1), warps 4-7 issue DFMA chains (mode & 2)
__global__ void __launch_bounds__(256) k_pipe_mix(int iters, int mode, double* sink) {
  const int warp = threadIdx.x >> 5;
  double s = 0.0;
  if (warp < 4) {
    if (mode & 1) {
      double c[8][2];
#pragma unroll
      for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
      double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    }
  } else if (mode & 2) {
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k] = threadIdx.x * 1e-3 + k;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-12;
    for (int it = 0; it < iters * 8; ++it) {     // 8 DFMA per lane ~ same pipe time as one DMMA if rates match
#pragma unroll
      for (int k = 0; k < 8; ++k) c[k] = fma(c[k], a, b);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k];
  }
  if (s == 12345.678) sink[0] = s;
}

// throughput of the kernel-function evaluation itself (entries / s)
__global__ void __launch_bounds__(256) k_kern_rate(int iters, int kid, double* sink) {
  double r2 = 1e-3 * threadIdx.x + 1e-6 * blockIdx.x, s = 0.0;
  for (int it = 0; it < iters; ++it) {
    double k, h;
    kern_eval(kid, r2, 0.5, k, h);
    s += k + h;
    r2 += 1e-4;
  }
  if (s == 12345.678) sink[0] = s;
}

// 128x128 supertile core: mode 0 = shared-memory resident operands (no global traffic),
// mode 1 = stream nk k-tiles per CTA from a private region (HBM), mode 2 = all CTAs read the same tiles (L2)
template <bool TA, bool TBm>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm2_bench(const double* tiles, long tiles_per_cta, int nk, int mode,
                                                             double* out) {
  extern __shared__ __align__(128) double smem[];
  Frag2 f;
  G2Pipe pipe;
  pipe.init();
  Acc2 acc;
  acc.zero();
  if (mode == 0) {
    for (int i = threadIdx.x; i < G2_STAGE_ELEMS; i += NTHREADS) smem[i] = 1e-3 * (i & 255);
    __syncthreads();
    for (int k = 0; k < 2 * nk; ++k) mma_half<TA, TBm>(acc, smem + f.ta * HALF_ELEMS, smem + (2 + f.tb) * HALF_ELEMS, f);
  } else {
    const double* base = tiles + (mode == 1 ? (long)blockIdx.x * tiles_per_cta * TILE_ELEMS : 0L);
    gemm2_pipeline<TA, TBm>(
        acc, smem, pipe, 0, nk,
        [&](int k, int t) { return base + ((long)(4 * k + t) % tiles_per_cta) * TILE_ELEMS; },
        [&](int k, int t) { return base + ((long)(4 * k + 2 + t) % tiles_per_cta) * TILE_ELEMS; }, f);
  }
  double s = 0.0;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) s += acc.c[mi][ni][0] + acc.c[mi][ni][1];
  if (s == 12345.678) out[0] = s;
}

// the 128x128 diagonal-block routine of the Cholesky panels in isolation: every CTA factorises `reps` times a
// diagonally dominant block it builds in shared memory; out[0..] receives scratch tiles
// stage stamps of diag_block_128 (block 0, thread 0, last repetition): 0 entry | 1 64x64 factor of block 0 | 2 its
// inverse sweep | 3 L00/X00 written | 4 L10 = C10 X00' | 5 C11 - L10 L10' + 64x64 factor of block 1 | 6 its inverse sweep |
// 7 L11/X11 written | 8 M = L10 X00 | 9 X10 = -X11 M issued
struct ClockProbe {
  long long* out;
  __device__ __forceinline__ void operator()(int stage) const {
    if (out && blockIdx.x == 0 && threadIdx.x == 0) out[stage] = clock64();
  }
};
__global__ void __launch_bounds__(NTHREADS, 1) k_diag_bench(int reps, double* scratch, int* fail, double* ld,
                                                            long long* stamps) {
  extern __shared__ __align__(128) double smem[];
  double* dg = smem + G2_SMEM_ELEMS + TILE_ELEMS;
  double* g = scratch + (long)blockIdx.x * 6 * TILE_ELEMS;
  for (int it = 0; it < reps; ++it) {
    __syncthreads();
    for (int t = threadIdx.x; t < TILE_ELEMS; t += NTHREADS) {
      const int r = t >> 6, cc = t & 63;
      const double off = 0.01 / (1.0 + abs(r - cc));
      smem[r * LDA + cc] = (r == cc) ? 2.0 : off;
      smem[DIAG_P1 + swz(r, cc)] = 0.005 / (1.0 + abs(r + 64 - cc));
      smem[DIAG_P2 + swz(r, cc)] = (r == cc) ? 2.0 : off;
    }
    __syncthreads();
    diag_block_128(smem, dg, true, 0, 1 << 30, fail + blockIdx.x, g, g + TILE_ELEMS, g + 2 * TILE_ELEMS,
                   g + 3 * TILE_ELEMS, g + 4 * TILE_ELEMS, g + 5 * TILE_ELEMS, ld + 2 * blockIdx.x, ClockProbe{stamps});
  }
}

// accuracy of the lean elementary functions against libdevice: max relative error over a sweep (mode 0: exp_neg on
// u in [0, 745), mode 1: sqrt_pos on [1e-36, 1e12)), encoded in out[0] via atomicMax on the bit pattern
__global__ void __launch_bounds__(256) k_elem_accuracy(int mode, int per_thread, unsigned long long* out) {
  const long tid = blockIdx.x * 256L + threadIdx.x, nthreads = gridDim.x * 256L;
  double worst = 0.0;
  for (int it = 0; it < per_thread; ++it) {
    const double frac = (double)(tid * per_thread + it) / (double)(nthreads * per_thread);
    double got, ref;
    if (mode == 0) {
      const double u = 745.0 * frac * frac;          // denser near 0
      got = exp_neg(u);
      ref = exp(-u);
      if (u > 700.0) { ref = 0.0; }
    } else {
      const double x = 1e-36 * exp(110.5 * frac);    // 1e-36 .. 1e12
      got = sqrt_pos(x);
      ref = sqrt(x);
    }
    const double err = (ref == 0.0) ? fabs(got) : fabs(got / ref - 1.0);
    worst = fmax(worst, err);
  }
  atomicMax(out, (unsigned long long)__double_as_longlong(worst));
}

// task streams with the 128x128 core (one CTA per SM): `tasks` GEMMs of nk k-tiles back to back, prologue and epilogue
// store included, like the real short-loop kernels.  (The 128x64 two-CTAs-per-SM variant measured against it in round 1
// -- +1.5-3 %, profiles/r01_microbench_2cta.json -- was not adopted and its code is no longer part of the library.)
template <bool TA, bool TBm>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm2_tasks_bench(const double* tiles, long tiles_per_cta, int nk,
                                                                   int tasks, double* out) {
  extern __shared__ __align__(128) double smem[];
  Frag2 f;
  G2Pipe pipe;
  pipe.init();
  const double* base = tiles + (long)blockIdx.x * tiles_per_cta * TILE_ELEMS;
  double* dst = out + (long)blockIdx.x * 4 * TILE_ELEMS;
  for (int t = 0; t < tasks; ++t) {
    Acc2 acc;
    acc.zero();
    gemm2_pipeline<TA, TBm>(
        acc, smem, pipe, 0, nk,
        [&](int k, int tt) { return base + ((long)(4 * k + tt + t) % tiles_per_cta) * TILE_ELEMS; },
        [&](int k, int tt) { return base + ((long)(4 * k + 2 + tt + t) % tiles_per_cta) * TILE_ELEMS; }, f);
    store_acc2(dst + (2 * f.ta + f.tb) * TILE_ELEMS, acc, f);
  }
}

// 64x64 core (gemm_core.cuh), same modes
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm1_bench(const double* tiles, long tiles_per_cta, int nk, int mode,
                                                             double* out) {
  extern __shared__ __align__(128) double smem[];
  FragCoord fc;
  Acc acc;
  acc.zero();
  if (mode == 0) {
    for (int i = threadIdx.x; i < 2 * TILE_ELEMS; i += NTHREADS) smem[i] = 1e-3 * (i & 255);
    __syncthreads();
    for (int k = 0; k < nk; ++k) mma_tile<false, false>(acc, smem, smem + TILE_ELEMS, fc);
  } else {
    const double* base = tiles + (mode == 1 ? (long)blockIdx.x * tiles_per_cta * TILE_ELEMS : 0L);
    gemm_pipeline<false, false>(
        acc, smem, 0, nk, [&](int k) { return base + ((long)(2 * k) % tiles_per_cta) * TILE_ELEMS; },
        [&](int k) { return base + ((long)(2 * k + 1) % tiles_per_cta) * TILE_ELEMS; }, fc);
  }
  double s = 0.0;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) s += acc.c[mi][ni][0] + acc.c[mi][ni][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace gpsat
