// gpsat_b200: 128x128 "supertile" FP64 GEMM core on the DMMA pipe (mma.sync m8n8k4 f64).
//
// One CTA (8 warps, 2 along M x 4 along N, 64x32 per warp) owns a 2x2 group of 64x64 output
// tiles and streams operand tiles (packed swizzled 32 KiB blobs, see common.cuh) through a
// 3-stage ring of 32-deep k-slices:  per slice 4 half-tiles (2 of A, 2 of B) = 64 KiB,
// 128x128x32 FMAs -> 32 KiB of operand traffic per 64^3 tile product (half of the 64x64 core).
// The ring is filled by TMA bulk copies (cp.async.bulk -> SASS UBLKCP) issued by ONE thread and
// tracked by one mbarrier per stage (expect_tx / complete_tx); the 256 compute threads issue no
// copy instructions at all.  (The previous per-thread cp.async version lost 14 % of the DMMA rate:
// 4096 LDGSTS per slice queue in front of the fragment LDS in the same MIO pipe.)
// Fragments are double-buffered in registers so the shared-memory loads of step k+4 are in
// flight while the 32 DMMAs of step k issue.
//
// Operand orientation (per tile, straight from the swizzled image):
//   TA  = false: A[m][k] = Atile(m, k)  (k along tile columns)   TA  = true: A[m][k] = Atile(k, m)
//   TBm = false: B[k][n] = Btile(n, k)                            TBm = true: B[k][n] = Btile(k, n)
// A k-slice of a "k along columns" tile is one 64 x 32 column half (contiguous 16 KiB, row stride 32);
// a k-slice of a "k along rows" tile is rows [32h, 32h+32) of both column halves (two contiguous
// 8 KiB pieces, kept side by side in shared memory: [half][32 rows][32 cols]).
#pragma once
#include "common.cuh"

namespace gpsat {

constexpr int G2_STAGES = 3;
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
  static constexpr bool kHalves = false;   // every warp owns all 64 rows of its slab
  int lane, warp, q, r, ta, tb, nb0;
  __device__ __forceinline__ int mi_begin() const { return 0; }     // 8-row groups this warp COMPUTES ...
  __device__ __forceinline__ int mi_end() const { return 8; }
  __device__ __forceinline__ int st_begin() const { return 0; }     // ... and STORES
  __device__ __forceinline__ int st_end() const { return 8; }
  __device__ __forceinline__ Frag2() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    q = lane >> 2;
    r = lane & 3;
    // Warp w runs on scheduler w % 4 (the DMMA pipe is per scheduler).  The two warps of a scheduler own
    // diagonally opposite tiles of the 2x2 group, (ta, tb) and (1 - ta, 1 - tb): whenever one A tile or one B
    // tile of a k-step is structurally zero, every scheduler keeps exactly one busy warp, so the skipped
    // products shorten the step instead of idling half of the pipes.
    ta = warp >> 2;                   // which A tile (row tile of the 2x2 group)
    tb = ((warp >> 1) & 1) ^ ta;      // which B tile (column tile of the 2x2 group)
    nb0 = (warp & 1) * 32;            // 32-column slab inside that tile
  }
  __device__ __forceinline__ int row(int mi) const { return 8 * mi + q; }             // within tile ta
  __device__ __forceinline__ int col(int ni) const { return nb0 + 8 * ni + 2 * r; }   // within tile tb (and col+1)
};

// Frag2 with row skipping: a warp whose tile is the LAST tile row of the matrix computes only rows 0-31 of its slab when
// rows 32-63 of that tile are padding (their results are structural zeros).  `last_tile` = index of the last tile
// row (nb - 1), `tile0` = tile row of ta = 0, `half_pad` = rows 32-63 of the last tile are all padding.
struct Frag2H : Frag2 {
  static constexpr bool kHalves = true;
  int half;
  __device__ __forceinline__ Frag2H(int tile0, int last_tile, bool half_pad) : Frag2() {
    half = (half_pad && tile0 + ta == last_tile) ? 1 : 0;
  }
  __device__ __forceinline__ int mi_begin() const { return 0; }
  __device__ __forceinline__ int mi_end() const { return half == 1 ? 4 : 8; }
  // the skipped rows are still stored (as the zeros the accumulators were cleared to): later kernels read whole tiles
  __device__ __forceinline__ int st_begin() const { return 0; }
  __device__ __forceinline__ int st_end() const { return 8; }
};

// Warp map for DIAGONAL supertiles (output = tiles (0,0), (1,0), (1,1); tile (0,1) is the mirror image of (1,0) and
// is not computed).  With the Frag2 map the two warps that own (0,1) would idle and the step would still take as
// long as a full supertile (schedulers 0 and 1 keep two busy warps each).  Here the three tiles' six 64x32 slabs
// are dealt so that EVERY scheduler gets three 32x32 blocks (a full slab + half a slab): a diagonal-supertile k-step
// takes 3/4 of the time of a full one.
//   scheduler 0: warp 0 = (0,0) cols  0-31 all rows   | warp 4 = (1,1) cols 32-63 rows  0-31
//   scheduler 1: warp 1 = (0,0) cols 32-63 all rows   | warp 5 = (1,1) cols 32-63 rows 32-63
//   scheduler 2: warp 2 = (1,0) cols  0-31 all rows   | warp 6 = (1,1) cols  0-31 rows  0-31
//   scheduler 3: warp 3 = (1,0) cols 32-63 all rows   | warp 7 = (1,1) cols  0-31 rows 32-63
struct Frag2D {
  static constexpr bool kHalves = true;
  int lane, warp, q, r, ta, tb, nb0, half;   // half: 0 = rows 0-63, 1 = rows 0-31 (mi 0..3), 2 = rows 32-63 (mi 4..7)
  __device__ __forceinline__ Frag2D() {
    lane = threadIdx.x & 31;
    warp = threadIdx.x >> 5;
    q = lane >> 2;
    r = lane & 3;
    if (warp < 4) {
      ta = warp >> 1;            // warps 0,1 -> tile row 0; warps 2,3 -> tile row 1
      tb = 0;
      nb0 = (warp & 1) * 32;
      half = 0;
    } else {
      ta = 1;
      tb = 1;
      nb0 = (warp < 6) ? 32 : 0;
      half = 1 + (warp & 1);
    }
  }
  __device__ __forceinline__ int mi_begin() const { return half == 2 ? 4 : 0; }
  __device__ __forceinline__ int mi_end() const { return half == 1 ? 4 : 8; }
  __device__ __forceinline__ int st_begin() const { return mi_begin(); }   // the other half belongs to another warp
  __device__ __forceinline__ int st_end() const { return mi_end(); }
  __device__ __forceinline__ int row(int mi) const { return 8 * mi + q; }
  __device__ __forceinline__ int col(int ni) const { return nb0 + 8 * ni + 2 * r; }
};

// one thread: bulk-copy one 32-deep k-slice (16 KiB) of a tile into a shared-memory half
template <bool KROWS>   // KROWS: k runs along tile rows; else along columns
__device__ __forceinline__ void bulk_half(double* smem_half, const double* gmem_tile, int kh, uint64_t* bar) {
  if (KROWS) {
    bulk_g2s(smem_half, gmem_tile + kh * 1024, 8192, bar);
    bulk_g2s(smem_half + 1024, gmem_tile + HALF_ELEMS + kh * 1024, 8192, bar);
  } else {
    bulk_g2s(smem_half, gmem_tile + kh * HALF_ELEMS, 16384, bar);
  }
}

// acc += op(A_half) * op(B_half) for this warp's 64x32 slab, 32-deep slice resident in shared memory
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `last_loads_done()` is invoked once, after the fragments of the LAST k-step have arrived in registers (this warp will
// not read the slice again) and before that step's 32 DMMAs are issued: the ring uses it to release the slot early.
// MI0 / MI1: the 8-row groups [MI0, MI1) of the slab this warp computes (0, 8 = the whole 64x32 slab).
template <bool TA, bool TBm, int MI0, int MI1, class FRAG, class HOOK>
__device__ __forceinline__ void mma_half_r(Acc2& acc, const double* __restrict__ As, const double* __restrict__ Bs,
                                           const FRAG& f, HOOK last_loads_done) {
  int aoff[8], boff[4];
  const int sq = (f.q & 3) << 2;
#pragma unroll
  for (int mi = MI0; mi < MI1; ++mi) {
    const int m = 8 * mi + f.q;
    aoff[mi] = TA ? ((m >> 5) * 1024 + f.r * 32 + ((m & 31) ^ (f.r << 2))) : (m * 32 + f.r);
  }
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    const int n = f.nb0 + 8 * ni + f.q;
    boff[ni] = TBm ? ((n >> 5) * 1024 + f.r * 32 + ((n & 31) ^ (f.r << 2))) : (n * 32 + f.r);
  }
  double a[2][8], b[2][4];
#pragma unroll
  for (int mi = MI0; mi < MI1; ++mi) a[0][mi] = As[aoff[mi] + (TA ? 0 : (0 ^ sq))];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) b[0][ni] = Bs[boff[ni] + (TBm ? 0 : (0 ^ sq))];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int cur = ks & 1, nxt = cur ^ 1;
    if (ks + 1 < 8) {
      const int kk = (ks + 1) * 4;
#pragma unroll
      for (int mi = MI0; mi < MI1; ++mi) a[nxt][mi] = As[aoff[mi] + (TA ? kk * 32 : (kk ^ sq))];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[nxt][ni] = Bs[boff[ni] + (TBm ? kk * 32 : (kk ^ sq))];
    }
    if (ks == 7) {
      // last step: DMMAs that between them read every fragment register first -- once they have issued, the
      // scoreboard guarantees that all of this warp's shared-memory loads of the slice have returned -- then the
      // hook, then the remaining DMMAs (which hide whatever latency the hook started)
#pragma unroll
      for (int mi = MI0; mi < MI1; ++mi)
        dmma884(acc.c[mi][mi & 3][0], acc.c[mi][mi & 3][1], a[cur][mi], b[cur][mi & 3]);
      last_loads_done();
#pragma unroll
      for (int mi = MI0; mi < MI1; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          if (ni != (mi & 3)) dmma884(acc.c[mi][ni][0], acc.c[mi][ni][1], a[cur][mi], b[cur][ni]);
    } else {
#pragma unroll
      for (int mi = MI0; mi < MI1; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc.c[mi][ni][0], acc.c[mi][ni][1], a[cur][mi], b[cur][ni]);
    }
  }
}
// rows: 0 = the whole slab, 1 = rows 0-31 (mi 0..3), 2 = rows 32-63 (mi 4..7); warp-uniform.  DYN = false compiles the
// full-slab variant only.
template <bool TA, bool TBm, bool DYN, class FRAG, class HOOK>
__device__ __forceinline__ void mma_half_sel(Acc2& acc, const double* __restrict__ As, const double* __restrict__ Bs,
                                             const FRAG& f, int rows, HOOK last_loads_done) {
  if constexpr (DYN) {
    if (rows == 0) mma_half_r<TA, TBm, 0, 8>(acc, As, Bs, f, last_loads_done);
    else if (rows == 1) mma_half_r<TA, TBm, 0, 4>(acc, As, Bs, f, last_loads_done);
    else mma_half_r<TA, TBm, 4, 8>(acc, As, Bs, f, last_loads_done);
  } else {
    mma_half_r<TA, TBm, 0, 8>(acc, As, Bs, f, last_loads_done);
  }
}
template <class FRAG>
__device__ __forceinline__ int own_rows(const FRAG& f) {
  if constexpr (FRAG::kHalves) return f.half;
  else return 0;
}
template <bool TA, bool TBm, class FRAG, class HOOK>
__device__ __forceinline__ void mma_half_h(Acc2& acc, const double* __restrict__ As, const double* __restrict__ Bs,
                                           const FRAG& f, HOOK last_loads_done) {
  mma_half_sel<TA, TBm, FRAG::kHalves>(acc, As, Bs, f, own_rows(f), last_loads_done);
}
template <bool TA, bool TBm, class FRAG>
__device__ __forceinline__ void mma_half(Acc2& acc, const double* __restrict__ As, const double* __restrict__ Bs,
                                         const FRAG& f) {
  mma_half_h<TA, TBm>(acc, As, Bs, f, NoHook());
}

// Ring state of one CTA: mbarriers in shared memory + the number of slices that went through the ring so far
// (stage and phase parity of slice n are n % G2_STAGES and (n / G2_STAGES) & 1), so several pipelines can run
// back to back in one kernel.  init() must be called by all threads once, before the first pipeline.
// done[s] counts the warps that have finished reading the slice in ring slot s.
struct G2Pipe {
  uint64_t* full;
  int* done;
  uint32_t count;
  __device__ __forceinline__ void init() {
    __shared__ __align__(8) uint64_t bars[G2_STAGES];
    __shared__ int cnt[G2_STAGES];
    full = bars;
    done = cnt;
    count = 0;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < G2_STAGES; ++s) {
        mbar_init(bars + s, 1);
        cnt[s] = 0;
      }
      fence_mbar_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ int stage(uint32_t n) const { return (int)(n % G2_STAGES); }
  __device__ __forceinline__ void wait(uint32_t n) const { mbar_wait(full + stage(n), (n / G2_STAGES) & 1u); }
};

struct NoTail {
  __device__ __forceinline__ const double* operator()(int, int) const { return nullptr; }
};
// Row selector of a pipeline: rows_of(k, kh, ta) says which rows of the A-side tile `ta` can receive a non-zero
// contribution from the 32-deep slice kh of k-tile k (0 = all, 1 = rows 0-31 only, 2 = rows 32-63 only) -- the slices
// of TRIANGULAR operand tiles that are half zeros.  The rest of the slab is not multiplied.
struct AllRows {
  static constexpr bool kAll = true;
  __device__ __forceinline__ int operator()(int, int, int) const { return 0; }
};
// the A-side tile `ta` is the lower-triangular DIAGONAL tile at k == d0 + ta: its slice `kh` reaches only `rows`
struct DiagRows {
  static constexpr bool kAll = false;
  int d0, kh, rows;
  __device__ __forceinline__ int operator()(int k, int kh_, int ta) const {
    return (k == d0 + ta && kh_ == kh) ? rows : 0;
  }
};

// acc += sum_{k = kbeg}^{kend-1} [A_k^0; A_k^1] * [B_k^0  B_k^1]  over 64-deep tile steps.
// a_of(k, t) / b_of(k, t), t in {0,1}: global pointer of the tile or nullptr (structurally zero /
// out of range: the copy and the products that would use it are skipped; nullness must be
// CTA-uniform).  smem: G2_SMEM_ELEMS doubles.
// There is no block barrier inside the loop: a warp waits only for the mbarrier of the slice it is about to read;
// when it is done with a slice it bumps the slot's counter, and the LAST of the 8 warps to do so issues the TMA
// copies of the slice that reuses the slot (so warps may drift up to two slices apart and the DMMA pipes never
// drain at slice boundaries).
// TAIL: after the last k-slice two more ring slots are filled with whole tiles tail_of(e, t), e, t in {0,1}
// (e.g. the C tiles an epilogue needs), fetched while the last slices are being multiplied; slot e lands in
// tail[e] (tile t at tail[e] + t * TILE_ELEMS, missing tiles are not touched).  Without TAIL the call ends with
// every copy consumed and a __syncthreads(); with TAIL the caller reads the tiles and then must __syncthreads()
// before the ring is reused.
template <bool TA, bool TBm, bool TAIL, class FA, class FB, class FRAG, class FT, class ROWS = AllRows>
__device__ __forceinline__ void gemm2_pipeline_t(Acc2& acc, double* smem, G2Pipe& p, int kbeg, int kend, FA a_of,
                                                 FB b_of, const FRAG& f, FT tail_of, const double** tail,
                                                 int drop_last = 0, ROWS rows_of = ROWS()) {
  // drop_last = 1: the second 32-deep half of the LAST k-tile holds only padding (structural zeros) and is not streamed
  const int nsl = (kend > kbeg) ? 2 * (kend - kbeg) - drop_last : 0;   // 32-deep slices
  const int ntot = nsl + (TAIL ? 2 : 0);
  if (ntot == 0) return;
  auto issue = [&](int sl) {    // one thread
    const uint32_t n = p.count + sl;
    double* st = smem + p.stage(n) * G2_STAGE_ELEMS;
    uint64_t* bar = p.full + p.stage(n);
    if (sl < nsl) {
      const int k = kbeg + (sl >> 1), kh = sl & 1;
      const double* pa0 = a_of(k, 0);
      const double* pa1 = a_of(k, 1);
      const double* pb0 = b_of(k, 0);
      const double* pb1 = b_of(k, 1);
      mbar_expect_tx(bar, 16384u * ((pa0 != nullptr) + (pa1 != nullptr) + (pb0 != nullptr) + (pb1 != nullptr)));
      if (pa0) bulk_half<TA>(st, pa0, kh, bar);
      if (pa1) bulk_half<TA>(st + HALF_ELEMS, pa1, kh, bar);
      if (pb0) bulk_half<TBm>(st + 2 * HALF_ELEMS, pb0, kh, bar);
      if (pb1) bulk_half<TBm>(st + 3 * HALF_ELEMS, pb1, kh, bar);
    } else {
      const int e = sl - nsl;
      const double* t0 = tail_of(e, 0);
      const double* t1 = tail_of(e, 1);
      mbar_expect_tx(bar, 32768u * ((t0 != nullptr) + (t1 != nullptr)));
      if (t0) bulk_g2s(st, t0, 32768, bar);
      if (t1) bulk_g2s(st + TILE_ELEMS, t1, 32768, bar);
    }
  };
  if (threadIdx.x == 0) {
    for (int sl = 0; sl < G2_STAGES && sl < ntot; ++sl) issue(sl);
  }
  for (int sl = 0; sl < nsl; ++sl) {
    const uint32_t n = p.count + sl;
    p.wait(n);               // the slice has landed
    const int k = kbeg + (sl >> 1);
    const bool refill = sl + G2_STAGES < ntot;
    int* cnt = p.done + p.stage(n);
    int seen = -1;                      // lane 0: the slot counter before this warp's release
    // rows of this warp's slab that the slice can contribute to: the warp's own share, cut by the operand's structure
    int rows = own_rows(f);
    bool none = false;
    if constexpr (!ROWS::kAll) {
      const int sel = rows_of(k, sl & 1, f.ta);
      if (sel != 0) {
        if (rows == 0) rows = sel;
        else if (rows != sel) none = true;
      }
    }
    if (!none && a_of(k, f.ta) != nullptr && b_of(k, f.tb) != nullptr) {
      const double* st = smem + p.stage(n) * G2_STAGE_ELEMS;
      // the slot is released from inside the last k-step (its operands are in registers by then), so the latency of
      // the shared-memory atomic hides behind that step's 32 DMMAs instead of idling the pipe at the slice boundary
      mma_half_sel<TA, TBm, (FRAG::kHalves || !ROWS::kAll)>(
          acc, st + f.ta * HALF_ELEMS, st + (2 + f.tb) * HALF_ELEMS, f, rows, [&]() {
            if (refill && f.lane == 0) seen = atomicAdd(cnt, 1);
          });
    } else if (refill && f.lane == 0) {
      seen = atomicAdd(cnt, 1);
    }
    if (seen == NTHREADS / 32 - 1) {    // last warp out refills the slot
      atomicExch(cnt, 0);
      __threadfence_block();
      issue(sl + G2_STAGES);
    }
  }
  if (TAIL) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      p.wait(p.count + nsl + e);
      tail[e] = smem + p.stage(p.count + nsl + e) * G2_STAGE_ELEMS;
    }
  } else {
    __syncthreads();
  }
  p.count += ntot;
}

template <bool TA, bool TBm, class FA, class FB, class FRAG, class ROWS = AllRows>
__device__ __forceinline__ void gemm2_pipeline(Acc2& acc, double* smem, G2Pipe& p, int kbeg, int kend, FA a_of,
                                               FB b_of, const FRAG& f, int drop_last = 0, ROWS rows_of = ROWS()) {
  gemm2_pipeline_t<TA, TBm, false>(acc, smem, p, kbeg, kend, a_of, b_of, f, NoTail(), nullptr, drop_last, rows_of);
}

// store this warp's 64x32 slab into a swizzled 64x64 tile (global or shared), scaled
template <class FRAG>
__device__ __forceinline__ void store_acc2(double* __restrict__ tile, const Acc2& acc, const FRAG& f,
                                           double scale = 1.0) {
  const int m0 = f.st_begin(), m1 = f.st_end();
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    if (mi < m0 || mi >= m1) continue;      // rows this warp does not own (Frag2D half slabs)
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
