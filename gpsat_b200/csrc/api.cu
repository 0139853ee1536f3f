// gpsat_b200: host orchestration + C ABI (see include/gpsat_b200.h).
#include "../../include/gpsat_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "engine.cuh"
#include "selection.cuh"
#include "microbench.cuh"
#include "sgpr.cuh"
#include "postproc.cuh"
#include "preproc.cuh"

using namespace gpsat;

static_assert(GPSAT_MAXD == MAXD && GPSAT_MAXP == MAXP, "header / device constants out of sync");
static_assert(sizeof(gpsat_sel_spec) == sizeof(SelSpec), "selection spec layout mismatch");

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return fail((int)e__, std::string(#call) + ": " + cudaGetErrorString(e__));         \
  } while (0)

constexpr int MAX_GROUPS = 8;
constexpr int MIN_GROUP_SLOTS = 32;
// Look-ahead Cholesky panels for rounds of at most this many 128-row supertile rows (N <= 767), with LA_GROUPS slot
// groups instead of 3.  Measured (profiles/r02_results.md): on c1 (N 400-640) the fused one-launch panels lose
// throughput as slot groups are added (CTAs spinning on a diagonal block hold an SM: 1657 / 1619 / 1587 experts/s at
// 3 / 6 / 8 groups) while the look-ahead panels gain (1722 / 1766 / 1764); c2 and c5 are unchanged; on c3 / c4 sizes
// the two modes are equal within noise and the fused mode stays (its single-stream phase timing is 2 % better).
constexpr int LA_MAX_NSR = 6;
constexpr int LA_GROUPS = 6;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct gpsat_handle {
  int device = 0;
  size_t budget = 0;
  int n_sm = 148;
  long long launches = 0;
  bool profiling = false;
  double ms[6] = {0, 0, 0, 0, 0, 0};   // potrf, trtri, lauum, other (finalize), build, trace
  double fl[3] = {0, 0, 0};
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  std::vector<double> ev_flops;  // N^3/3 summed over active slots for each recorded round
  Buf Lt, Xt, Kt, quad, abuf, coords, yobs, ints, theta, logdet, gpart, fout, gout, states, order, pslot, pres, scratch, items;
  // sparse GPR: second tile pool (B = I + beta A'A'^T) and rectangular work matrices
  Buf Lt2, Xt2, Kt2, quad2, logdet2, fail2, sg_mm, sg_mn, sg_vec, sg_scal, sg_gpart, sg_ints, sg_beta, sg_ymean, sg_zeros;
  int* host_ints = nullptr;  // pinned
  size_t host_ints_cap = 0;
  int max_slots = 0;         // 0: 4 slots per SM (GPSAT_MAX_SLOTS overrides)
  int n_groups = 3;          // slot groups / streams of the optimiser (GPSAT_GROUPS overrides)
  bool n_groups_env = false; // GPSAT_GROUPS was given: no automatic choice
  cudaStream_t gstream[8] = {};
  cudaEvent_t gevent[8] = {};
  cudaEvent_t gevent2[8][2] = {};   // census events of the optimiser rounds, [group][parity]
  cudaEvent_t ev_fork = nullptr;
  int groups_ready = 0;
  bool attrs_set = false;
  int plan_slots = 0, plan_nbmax = 0;   // last make_plan (gpsat_last_plan)
  size_t plan_bytes_per_slot = 0;
  bool safe_panel = false;       // Cholesky panels as two launches (GPSAT_SAFE_PANEL=1, or after a flag-wait timeout)
  int la_max_nsr = LA_MAX_NSR;   // look-ahead panels for rounds of at most this many supertile rows (GPSAT_PANEL_LA)
  int* timeouts_dev = nullptr;   // [1] flag-wait timeouts of k_potrf_panel (device counter, see GPSAT_ESYNC)
  long long timeouts_seen = 0;   // value already reported to the caller
};

static int ensure(Buf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    e = cudaMalloc(&b.p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) return fail(GPSAT_ENOMEM, "cudaMalloc failed for " + std::to_string(bytes) + " bytes");
  b.cap = want;
  return 0;
}
#define ENS(buf, bytes)                     \
  do {                                      \
    int r__ = ensure(buf, (size_t)(bytes)); \
    if (r__) return r__;                    \
  } while (0)

extern "C" const char* gpsat_last_error(void) { return g_err.c_str(); }
extern "C" int gpsat_version(void) { return 100; }

extern "C" void gpsat_default_opts(gpsat_opt_options* o) {
  o->maxcor = 10;
  o->maxiter = 10000;
  o->maxfun = 15000;
  o->maxls = 20;
  o->ftol = 2.220446049250313e-09;
  o->gtol = 1e-5;
}

extern "C" int gpsat_create(gpsat_handle** out, int device, size_t mem_budget_bytes) {
  if (!out) return fail(GPSAT_EINVAL, "out is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GPSAT_ENOGPU, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(GPSAT_EINVAL, "bad device index");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(GPSAT_ENOGPU, std::string("gpsat_b200 is built for sm_100a only; found ") + prop.name);
  gpsat_handle* h = new gpsat_handle();
  h->device = device;
  h->n_sm = prop.multiProcessorCount;
  if (const char* eg = getenv("GPSAT_GROUPS")) {
    h->n_groups = std::max(1, std::min(MAX_GROUPS, atoi(eg)));
    h->n_groups_env = true;
  }
  if (const char* es = getenv("GPSAT_MAX_SLOTS")) h->max_slots = std::max(1, atoi(es));
  if (const char* sp = getenv("GPSAT_SAFE_PANEL")) h->safe_panel = atoi(sp) != 0;
  if (const char* la = getenv("GPSAT_PANEL_LA")) h->la_max_nsr = std::max(0, atoi(la));   // 0: never
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  h->budget = mem_budget_bytes ? mem_budget_bytes : (size_t)(0.7 * (double)free_b);
  CK(cudaFuncSetAttribute(k_potrf_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_trtri_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_trtri_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_lauum2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_predict2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_predict2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_pred_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_tgemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_tgemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  CK(cudaFuncSetAttribute(k_tgemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  // function attributes are per device: set here for this handle's device, not once per process
  CK(cudaFuncSetAttribute(k_select_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * sizeof(int)));
  CK(cudaMalloc(&h->timeouts_dev, sizeof(int)));
  CK(cudaMemset(h->timeouts_dev, 0, sizeof(int)));
  *out = h;
  return 0;
}

extern "C" int gpsat_destroy(gpsat_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  Buf* bs[] = {&h->Lt, &h->Xt, &h->Kt, &h->quad, &h->abuf, &h->coords, &h->yobs, &h->ints, &h->theta, &h->logdet, &h->gpart,
               &h->fout, &h->gout, &h->states, &h->order, &h->pslot, &h->pres, &h->scratch, &h->items,
               &h->Lt2, &h->Xt2, &h->Kt2, &h->quad2, &h->logdet2, &h->fail2, &h->sg_mm, &h->sg_mn, &h->sg_vec,
               &h->sg_scal, &h->sg_gpart, &h->sg_ints, &h->sg_beta, &h->sg_ymean, &h->sg_zeros};
  for (Buf* b : bs)
    if (b->p) cudaFree(b->p);
  if (h->host_ints) cudaFreeHost(h->host_ints);
  if (h->timeouts_dev) cudaFree(h->timeouts_dev);
  for (int g = 0; g < h->groups_ready; ++g) {
    cudaStreamDestroy(h->gstream[g]); cudaEventDestroy(h->gevent[g]);
    cudaEventDestroy(h->gevent2[g][0]); cudaEventDestroy(h->gevent2[g][1]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (auto e : h->ev) cudaEventDestroy(e);
  delete h;
  return 0;
}

extern "C" long long gpsat_launch_count(const gpsat_handle* h) { return h ? h->launches : 0; }
extern "C" int gpsat_last_plan(const gpsat_handle* h, int* slots, int* nbmax, size_t* bytes_per_slot,
                               size_t* budget_bytes) {
  if (!h) return GPSAT_EINVAL;
  if (slots) *slots = h->plan_slots;
  if (nbmax) *nbmax = h->plan_nbmax;
  if (bytes_per_slot) *bytes_per_slot = h->plan_bytes_per_slot;
  if (budget_bytes) *budget_bytes = h->budget;
  return 0;
}
extern "C" long long gpsat_sync_timeouts(gpsat_handle* h) {
  if (!h || !h->timeouts_dev) return 0;
  int v = 0;
  cudaSetDevice(h->device);
  if (cudaMemcpy(&v, h->timeouts_dev, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}
// after a call's final synchronisation: new flag-wait timeouts since the last report -> GPSAT_ESYNC
static int check_sync_timeouts(gpsat_handle* h) {
  const long long v = gpsat_sync_timeouts(h);
  if (v > h->timeouts_seen) {
    const long long d = v - h->timeouts_seen;
    h->timeouts_seen = v;
    h->safe_panel = true;      // in-order CTA dispatch did not hold on this device / under this tool: stop relying on it
    return fail(GPSAT_ESYNC, std::to_string(d) + " Cholesky panel CTA(s) timed out waiting for a diagonal block; the "
                             "affected evaluations were treated as non-positive-definite (f = +inf)");
  }
  return 0;
}
extern "C" int gpsat_set_profiling(gpsat_handle* h, int enabled) {
  if (!h) return GPSAT_EINVAL;
  h->profiling = enabled != 0;
  for (int k = 0; k < 6; ++k) h->ms[k] = 0;
  for (int k = 0; k < 3; ++k) h->fl[k] = 0;
  return 0;
}
extern "C" int gpsat_get_profile(gpsat_handle* h, double* a, double* b, double* c, double* d, double* fa,
                                 double* fb, double* fc, double* ms_build, double* ms_trace) {
  if (!h) return GPSAT_EINVAL;
  *a = h->ms[0]; *b = h->ms[1]; *c = h->ms[2]; *d = h->ms[3];
  *fa = h->fl[0]; *fb = h->fl[1]; *fc = h->fl[2];
  if (ms_build) *ms_build = h->ms[4];
  if (ms_trace) *ms_trace = h->ms[5];
  return 0;
}

// ------------------------------------------------------------------------------------------
// slot pool planning and workspace
// ------------------------------------------------------------------------------------------
struct Plan {
  int S, nbmax, npmax, ntmax;
  std::vector<int> order;  // expert ids, descending n
};

static size_t slot_bytes(int nbmax) {
  const size_t ntmax = (size_t)nbmax * (nbmax + 1) / 2;
  const size_t npmax = (size_t)nbmax * TB;
  return 3 * ntmax * TILE_BYTES + (MAXD + 1) * npmax * 8 + nbmax * 8 + ntmax * NG * 8 + sizeof(LbfgsState) + 256;
}

static int make_plan(gpsat_handle* h, const gpsat_batch* b, Plan& pl, size_t extra_per_slot = 0) {
  const int E = b->n_experts;
  if (E <= 0) return fail(GPSAT_EINVAL, "n_experts must be positive");
  if (b->D < 1 || b->D > MAXD) return fail(GPSAT_EINVAL, "D out of range");
  pl.order.resize(E);
  std::iota(pl.order.begin(), pl.order.end(), 0);
  const long long* off = b->offsets_host;
  std::stable_sort(pl.order.begin(), pl.order.end(),
                   [&](int x, int y) { return (off[x + 1] - off[x]) > (off[y + 1] - off[y]); });
  long long nmax = off[pl.order[0] + 1] - off[pl.order[0]];
  long long nmin = off[pl.order[E - 1] + 1] - off[pl.order[E - 1]];
  if (nmin < 1) return fail(GPSAT_EINVAL, "every expert needs at least one observation");
  pl.nbmax = (int)(nmax / TB) + 1;
  pl.npmax = pl.nbmax * TB;
  pl.ntmax = pl.nbmax * (pl.nbmax + 1) / 2;
  const size_t per = slot_bytes(pl.nbmax) + extra_per_slot;
  long long cap = (long long)(h->budget / per);
  if (cap < 1) return fail(GPSAT_ENOMEM, "one expert of this size does not fit the memory budget");
  pl.S = (int)std::min<long long>(std::min<long long>(E, cap), h->max_slots > 0 ? h->max_slots : 4LL * h->n_sm);
  h->plan_slots = pl.S; h->plan_nbmax = pl.nbmax; h->plan_bytes_per_slot = per;
  return 0;
}

struct Work {
  SlotCtx c;
  SlotAux a;
};

// the sub-pool [s0, s0 + Sg) of a slot pool
static SlotCtx slot_view(const SlotCtx& c, int s0, int Sg) {
  SlotCtx v = c;
  v.S = Sg;
  v.Lt += (size_t)s0 * c.tile_stride; v.Xt += (size_t)s0 * c.tile_stride; v.Kt += (size_t)s0 * c.tile_stride;
  v.coords += (size_t)s0 * MAXD * c.npmax; v.yobs += (size_t)s0 * c.npmax;
  v.n += s0; v.nb += s0; v.active += s0; v.fail += s0; v.pflag += s0;
  v.theta += (size_t)s0 * MAXP; v.logdet_part += (size_t)s0 * c.nbmax; v.quad += s0;
  v.gpart += (size_t)s0 * c.ntmax * NG; v.fout += s0; v.gout += (size_t)s0 * MAXP;
  return v;
}

static int ensure_groups(gpsat_handle* h, int G, int S) {
  if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  for (int g = h->groups_ready; g < G; ++g) {
    CK(cudaStreamCreateWithFlags(&h->gstream[g], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->gevent[g], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->gevent2[g][0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->gevent2[g][1], cudaEventDisableTiming));
    h->groups_ready = g + 1;
  }
  const size_t need = (size_t)MAX_GROUPS * 2 * 3 * S + 8;
  if (h->host_ints_cap < need) {
    if (h->host_ints) cudaFreeHost(h->host_ints);
    h->host_ints = nullptr;
    CK(cudaMallocHost(&h->host_ints, need * sizeof(int)));
    h->host_ints_cap = need;
  }
  return 0;
}

static int setup_work(gpsat_handle* h, const gpsat_batch* b, const Plan& pl, Work& w, cudaStream_t st) {
  const int S = pl.S;
  ENS(h->Lt, (size_t)S * pl.ntmax * TILE_BYTES);
  ENS(h->Xt, (size_t)S * pl.ntmax * TILE_BYTES);
  ENS(h->Kt, (size_t)S * pl.ntmax * TILE_BYTES);
  ENS(h->quad, (size_t)S * 8);
  ENS(h->coords, (size_t)S * MAXD * pl.npmax * 8);
  ENS(h->yobs, (size_t)S * pl.npmax * 8);
  ENS(h->ints, (size_t)(7 * S + 8) * sizeof(int));
  ENS(h->theta, (size_t)S * MAXP * 8);
  ENS(h->logdet, (size_t)S * pl.nbmax * 8);
  ENS(h->gpart, (size_t)S * pl.ntmax * NG * 8);
  ENS(h->fout, (size_t)S * 8);
  ENS(h->gout, (size_t)S * MAXP * 8);
  ENS(h->states, (size_t)S * sizeof(LbfgsState));
  ENS(h->order, (size_t)b->n_experts * sizeof(int));
  if (h->host_ints_cap < (size_t)(3 * S + 8)) {
    if (h->host_ints) cudaFreeHost(h->host_ints);
    CK(cudaMallocHost(&h->host_ints, (size_t)(3 * S + 8) * sizeof(int)));
    h->host_ints_cap = 3 * S + 8;
  }
  int* ip = (int*)h->ints.p;
  SlotCtx& c = w.c;
  c.nvar_override = -1.0;
  c.S = S; c.D = b->D; c.kid = b->kernel_id; c.nbmax = pl.nbmax; c.npmax = pl.npmax; c.ntmax = pl.ntmax;
  c.tile_stride = (long)pl.ntmax * TILE_ELEMS;
  c.Lt = (double*)h->Lt.p; c.Xt = (double*)h->Xt.p; c.Kt = (double*)h->Kt.p; c.quad = (double*)h->quad.p;
  c.coords = (double*)h->coords.p; c.yobs = (double*)h->yobs.p;
  c.n = ip; c.nb = ip + S; c.active = ip + 2 * S; c.fail = ip + 3 * S; c.pflag = ip + 6 * S;
  c.theta = (double*)h->theta.p; c.logdet_part = (double*)h->logdet.p; c.gpart = (double*)h->gpart.p;
  c.fout = (double*)h->fout.p; c.gout = (double*)h->gout.p;
  c.timeouts = h->timeouts_dev;
  w.a.slot_expert = ip + 4 * S;
  w.a.queue_head = ip + 5 * S;
  w.a.states = (LbfgsState*)h->states.p;
  CK(cudaMemsetAsync(h->ints.p, 0, (size_t)(7 * S + 8) * sizeof(int), st));
  CK(cudaMemcpyAsync(h->order.p, pl.order.data(), (size_t)b->n_experts * sizeof(int), cudaMemcpyHostToDevice, st));
  return 0;
}

static BatchIn make_batch_in(const gpsat_batch* b, const double* theta_dev, const int* order_dev) {
  BatchIn bi;
  bi.E = b->n_experts; bi.D = b->D;
  bi.coords = b->coords_dev; bi.obs = b->obs_dev; bi.offsets = b->offsets_dev;
  bi.order = order_dev; bi.theta0 = theta_dev;
  for (int d = 0; d < MAXD; ++d) bi.coords_scale[d] = (d < b->D && b->coords_scale[d] != 0.0) ? b->coords_scale[d] : 1.0;
  bi.obs_scale = (b->obs_scale != 0.0) ? b->obs_scale : 1.0;
  bi.obs_mean_local = b->obs_mean_local;
  bi.obs_mean_out = b->obs_mean_out_dev;
  return bi;
}

static cudaEvent_t next_event(gpsat_handle* h) {
  if (h->ev_used == h->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev.push_back(e);
  }
  return h->ev[h->ev_used++];
}

// one objective evaluation for all active slots.  nbm: max 64-blocks among active slots.
enum { RR_BUILD = 1, RR_INVERSE = 2, RR_LAUUM = 4, RR_TRACE = 8, RR_FINALIZE = 16, RR_PROFILE = 32 };
static int run_round_flags(gpsat_handle* h, const SlotCtx& c, int nbm, int flags, cudaStream_t st,
                           double flops_third) {
  const bool prof = h->profiling && (flags & RR_PROFILE);
  const int nsr = (nbm + 1) / 2, ntm = nbm * (nbm + 1) / 2;

  if (prof) { cudaEventRecord(next_event(h), st); h->ev_flops.push_back(flops_third); }
  if (flags & RR_BUILD) {
    switch (c.kid) {     // kernel family resolved at compile time inside the elementwise kernels
      case K_RBF: k_build<K_RBF><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      case K_MATERN32: k_build<K_MATERN32><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      case K_MATERN52: k_build<K_MATERN52><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      default: k_build<K_MATERN12><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
    }
    ++h->launches;
  }
  if (prof) cudaEventRecord(next_event(h), st);
  if (!h->safe_panel && nsr >= 2 && nsr <= h->la_max_nsr) {
    // look-ahead mode (small matrices): panel 0's diagonal blocks, then one launch per panel J holding its
    // off-diagonal blocks, whose row J + 1 CTAs go on to factorise the diagonal block of panel J + 1 (gpr2.cuh)
    k_potrf_panel<<<c.S, NTHREADS, PANEL_SMEM_BYTES, st>>>(c, 0, nsr, 0, 0);
    ++h->launches;
    for (int J = 0; J + 1 < nsr; ++J) {
      k_potrf_panel<<<c.S * (nsr - J - 1), NTHREADS, PANEL_SMEM_BYTES, st>>>(c, J, nsr, c.S, 1);
      ++h->launches;
    }
  } else {
    for (int J = 0; J < nsr; ++J) {
      const int n_off = c.S * (nsr - J - 1);
      if (!h->safe_panel) {
        k_potrf_panel<<<c.S + n_off, NTHREADS, PANEL_SMEM_BYTES, st>>>(c, J, nsr, 0, 0);
        ++h->launches;
      } else {      // diagonal blocks complete (kernel boundary) before any block that waits for their flag starts
        k_potrf_panel<<<c.S, NTHREADS, PANEL_SMEM_BYTES, st>>>(c, J, nsr, 0, 0);
        ++h->launches;
        if (n_off > 0) {
          k_potrf_panel<<<n_off, NTHREADS, PANEL_SMEM_BYTES, st>>>(c, J, nsr, c.S, 0);
          ++h->launches;
        }
      }
    }
  }
  k_quad<<<c.S, NTHREADS, 0, st>>>(c);
  ++h->launches;
  if (prof) cudaEventRecord(next_event(h), st);
  if (flags & RR_INVERSE) {
    for (int hh = 1; hh < nsr; hh *= 2) {
      const int nblk = (nsr + 2 * hh - 1) / (2 * hh);
      k_trtri_pass1<<<dim3(nblk * hh * hh, c.S), NTHREADS, SMEM2_BYTES, st>>>(c, hh);
      k_trtri_pass2<<<dim3(nblk * hh * hh, c.S), NTHREADS, SMEM2_BYTES, st>>>(c, hh);
      h->launches += 2;
    }
  }
  if (prof) cudaEventRecord(next_event(h), st);
  if (flags & RR_LAUUM) {
    k_lauum2<<<dim3(nsr * (nsr + 1) / 2, c.S), NTHREADS, SMEM2_BYTES, st>>>(c);
    ++h->launches;
  }
  if (prof) cudaEventRecord(next_event(h), st);
  if (flags & RR_TRACE) {
    switch (c.kid) {
      case K_RBF: k_grad_trace<K_RBF><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      case K_MATERN32: k_grad_trace<K_MATERN32><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      case K_MATERN52: k_grad_trace<K_MATERN52><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
      default: k_grad_trace<K_MATERN12><<<dim3(ntm, c.S), 256, 0, st>>>(c); break;
    }
    ++h->launches;
  }
  if (prof) cudaEventRecord(next_event(h), st);
  if (flags & RR_FINALIZE) {
    k_finalize2<<<c.S, NTHREADS, 0, st>>>(c, (flags & RR_TRACE) ? 1 : 0);
    ++h->launches;
  }
  if (prof) cudaEventRecord(next_event(h), st);
  CK(cudaGetLastError());
  return 0;
}
static int run_round(gpsat_handle* h, const SlotCtx& c, int nbm, bool inverse, bool grad, cudaStream_t st,
                     double flops_third) {
  return run_round_flags(h, c, nbm, RR_BUILD | RR_FINALIZE | RR_PROFILE | (inverse ? RR_INVERSE : 0) |
                                        (grad ? (RR_LAUUM | RR_TRACE) : 0), st, flops_third);
}

static void harvest_profile(gpsat_handle* h, bool inverse, bool grad) {
  if (!h->profiling) { h->ev_used = 0; h->ev_flops.clear(); return; }
  const size_t rounds = h->ev_used / 7;
  for (size_t r = 0; r < rounds; ++r) {
    float t[6];   // build | potrf + quad | trtri | lauum | trace | finalize
    for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], h->ev[7 * r + k], h->ev[7 * r + k + 1]);
    h->ms[4] += t[0]; h->ms[0] += t[1]; h->ms[1] += t[2]; h->ms[2] += t[3]; h->ms[5] += t[4]; h->ms[3] += t[5];
    h->fl[0] += h->ev_flops[r];
    if (inverse) h->fl[1] += h->ev_flops[r];
    if (grad) h->fl[2] += h->ev_flops[r];
  }
  h->ev_used = 0;
  h->ev_flops.clear();
}

static int nb_of(const long long* off, int e) { return (int)((off[e + 1] - off[e]) / TB) + 1; }
static double cube3(const long long* off, int e) {
  const double n = (double)(off[e + 1] - off[e]);
  return n * n * n / 3.0;
}

static TransformSpec identity_transforms(int D) {
  TransformSpec tr;
  memset(&tr, 0, sizeof(tr));
  tr.np = D + 2;
  tr.nfree = 0;
  return tr;
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
extern "C" int gpsat_gpr_eval(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev, double* f_dev,
                              double* grad_dev, void* stream) {
  if (!h || !b || !theta_dev) return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  Plan pl;
  int r = make_plan(h, b, pl);
  if (r) return r;
  Work w;
  r = setup_work(h, b, pl, w, st);
  if (r) return r;
  BatchIn bi = make_batch_in(b, theta_dev, (const int*)h->order.p);
  TransformSpec tr = identity_transforms(b->D);
  const bool grad = grad_dev != nullptr;
  for (int first = 0; first < b->n_experts; first += pl.S) {
    const int count = std::min(pl.S, b->n_experts - first);
    k_slot_init<<<pl.S, NTHREADS, 0, st>>>(w.c, w.a, bi, tr, first, count, 0);
    ++h->launches;
    double fl = 0;
    for (int k = 0; k < count; ++k) fl += cube3(b->offsets_host, pl.order[first + k]);
    r = run_round(h, w.c, nb_of(b->offsets_host, pl.order[first]), grad, grad, st, fl);
    if (r) return r;
    k_eval_scatter<<<(count + 127) / 128, 128, 0, st>>>(w.c, w.a, count, f_dev, grad_dev, b->D + 2);
    ++h->launches;
  }
  CK(cudaStreamSynchronize(st));
  harvest_profile(h, grad, grad);
  CK(cudaGetLastError());
  return check_sync_timeouts(h);
}

// ------------------------------------------------------------------------------------------
// optimisation
// ------------------------------------------------------------------------------------------
extern "C" int gpsat_gpr_optimise(gpsat_handle* h, const gpsat_batch* b, const double* theta0_dev,
                                  const gpsat_transforms* trs, const gpsat_opt_options* opts,
                                  double* theta_out_dev, double* fobj_out_dev, int* status_out_dev,
                                  int* nit_out_dev, int* nfev_out_dev, void* stream) {
  if (!h || !b || !theta0_dev || !trs || !theta_out_dev || !fobj_out_dev || !status_out_dev || !nit_out_dev ||
      !nfev_out_dev)
    return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  gpsat_opt_options od;
  gpsat_default_opts(&od);
  if (opts) od = *opts;
  if (od.maxcor < 1 || od.maxcor > LB_M) return fail(GPSAT_EINVAL, "maxcor must be in [1, 10]");
  Plan pl;
  int r = make_plan(h, b, pl);
  if (r) return r;
  Work w;
  r = setup_work(h, b, pl, w, st);
  if (r) return r;
  BatchIn bi = make_batch_in(b, theta0_dev, (const int*)h->order.p);
  TransformSpec tr;
  memset(&tr, 0, sizeof(tr));
  tr.np = b->D + 2;
  for (int p = 0; p < tr.np; ++p) {
    tr.kind[p] = trs->kind[p];
    tr.low[p] = trs->low[p];
    tr.high[p] = trs->high[p];
    if (trs->trainable[p]) tr.free_idx[tr.nfree++] = p;
  }
  if (tr.nfree == 0) return fail(GPSAT_EINVAL, "no trainable parameters (use gpsat_gpr_eval)");
  LbfgsOpts lo;
  lo.m = od.maxcor; lo.maxiter = od.maxiter; lo.maxfun = od.maxfun; lo.maxls = od.maxls;
  lo.factr = od.ftol / 2.220446049250313e-16;
  lo.pgtol = od.gtol;
  OptOut out{theta_out_dev, fobj_out_dev, status_out_dev, nit_out_dev, nfev_out_dev};
  const int S = pl.S;
  const int count = std::min(S, b->n_experts);
  k_slot_init<<<S, NTHREADS, 0, st>>>(w.c, w.a, bi, tr, 0, count, 1);
  ++h->launches;
  CK(cudaMemcpyAsync(w.a.queue_head, &count, sizeof(int), cudaMemcpyHostToDevice, st));
  // The slot pool is split into independent groups, each advancing its own rounds on its own stream
  // (all groups pull new experts from the one device work queue).  While one group sits in a serial or
  // latency-bound stretch (diagonal-block factorisations, launch tails, the optimiser step, the host's
  // look at the slot table) the other groups' CTAs fill the idle SMs.  Phase profiling keeps one group
  // so that the event timings are not overlapped.
  // (small matrices run the look-ahead panels, which keep gaining from more groups: see LA_MAX_NSR)
  const bool la_batch = !h->safe_panel && (pl.nbmax + 1) / 2 <= h->la_max_nsr;
  const int want_groups = (la_batch && !h->n_groups_env) ? LA_GROUPS : h->n_groups;
  int G = (h->profiling || S < 2 * MIN_GROUP_SLOTS) ? 1 : std::min(want_groups, S / MIN_GROUP_SLOTS);
  G = std::max(1, std::min(G, MAX_GROUPS));
  r = ensure_groups(h, G, S);
  if (r) return r;
  // The host only needs two numbers per group and round -- are there active slots, and the largest matrix among
  // them (grid sizes) -- and both may be STALE: slots are refilled from a queue sorted by descending size, so the
  // largest active matrix never grows, and a round launched for a group that has meanwhile finished is a handful of
  // early-exit grids.  So the census that sizes round r is the one taken after round r - 1 - LOOKAHEAD: the host
  // queues round r while round r - 1 is still running and the device never waits for the host between rounds
  // (GPSAT_LOOKAHEAD=0 restores the lock-step loop; phase profiling uses it so that the per-round flop counts are exact).
  struct Group { SlotCtx c; SlotAux a; int s0, Sg; bool done; long long rounds; int nact, nbm; double fl; };
  Group grp[MAX_GROUPS];
  int lookahead = h->profiling ? 0 : 1;
  if (const char* el = getenv("GPSAT_LOOKAHEAD")) lookahead = atoi(el) > 0 ? 1 : 0;
  if (h->profiling) lookahead = 0;
  CK(cudaEventRecord(h->ev_fork, st));
  for (int g = 0; g < G; ++g) {
    Group& q = grp[g];
    q.s0 = (int)((long long)S * g / G);
    q.Sg = (int)((long long)S * (g + 1) / G) - q.s0;
    q.c = slot_view(w.c, q.s0, q.Sg);
    q.a = w.a;
    q.a.slot_expert += q.s0;
    q.a.states += q.s0;
    q.done = false;
    q.rounds = 0;
    q.nact = q.nbm = 0;
    q.fl = 0;
    cudaStream_t sg = (G == 1) ? st : h->gstream[g];
    if (G > 1) CK(cudaStreamWaitEvent(sg, h->ev_fork, 0));
  }
  const long long max_rounds = (long long)(b->n_experts / std::max(1, S / G) + 2) * (od.maxfun + od.maxls + 2);
  int live = G;
  // GPSAT_TRACE=<file>: per group and round, host time at which the round was queued + the census it was sized by
  // (diagnostic: slot-pool utilisation over a batch, host gaps)
  FILE* trace = nullptr;
  if (const char* tp = getenv("GPSAT_TRACE")) trace = fopen(tp, "a");
  const auto t_trace0 = std::chrono::steady_clock::now();
  if (trace) fprintf(trace, "# optimise E=%d S=%d G=%d lookahead=%d\n", b->n_experts, S, G, lookahead);
  // census buffers: [group][parity][3 * Sg] ints (n | nb | active), events [group][parity]
  auto census_buf = [&](int g, long long c) { return h->host_ints + ((size_t)g * 2 + (size_t)(c & 1)) * 3 * S; };
  auto queue_census = [&](int g, long long c, cudaStream_t sg) -> int {
    Group& q = grp[g];
    int* hi = census_buf(g, c);
    for (int k = 0; k < 3; ++k)
      CK(cudaMemcpyAsync(hi + k * q.Sg, w.c.n + (size_t)k * S + q.s0, (size_t)q.Sg * sizeof(int),
                         cudaMemcpyDeviceToHost, sg));
    CK(cudaEventRecord(h->gevent2[g][c & 1], sg));
    return 0;
  };
  auto read_census = [&](int g, long long c) -> int {
    Group& q = grp[g];
    CK(cudaEventSynchronize(h->gevent2[g][c & 1]));
    const int* hi = census_buf(g, c);
    q.nact = 0; q.nbm = 0; q.fl = 0;
    for (int k = 0; k < q.Sg; ++k) {
      if (hi[2 * q.Sg + k]) {
        ++q.nact;
        q.nbm = std::max(q.nbm, hi[q.Sg + k]);
        const double n = (double)hi[k];
        q.fl += n * n * n / 3.0;
      }
    }
    return 0;
  };
  // The host never blocks on ONE group's census while another group's stream could be fed: a census that is not
  // ready yet is skipped in this pass (cudaEventQuery); only when no group made progress does the host wait -- on the
  // oldest outstanding census, which is the next thing that can unblock anything.
  bool limit_hit = false;
  while (live > 0 && !limit_hit) {
    bool progressed = false;
    int wait_g = -1;
    long long wait_c = 0;
    for (int g = 0; g < G; ++g) {
      Group& q = grp[g];
      if (q.done) continue;
      if (q.rounds >= max_rounds) { limit_hit = true; break; }
      cudaStream_t sg = (G == 1) ? st : h->gstream[g];
      // census index -1 = the slot table after k_slot_init; census c >= 0 = the table after round c
      if (q.rounds == 0) {
        r = queue_census(g, -1 + 2, sg);      // parity slot of "-1" (kept distinct from census 0)
        if (r) return r;
        r = read_census(g, -1 + 2);
        if (r) return r;
      } else {
        const long long c = q.rounds - 1 - lookahead;
        if (c >= 0) {
          if (G > 1 && cudaEventQuery(h->gevent2[g][c & 1]) == cudaErrorNotReady) {
            if (wait_g < 0) { wait_g = g; wait_c = c; }
            continue;                         // feed the other groups first
          }
          r = read_census(g, c);
          if (r) return r;
        }
      }
      progressed = true;
      if (trace) {
        // dry = 1: the group's previous round had already finished when this one was queued, i.e. its stream ran out
        // of work and waited for the host (with the one-round look-ahead this should be rare)
        const int dry = (q.rounds > 0 &&
                         cudaEventQuery(h->gevent2[g][(q.rounds - 1) & 1]) == cudaSuccess) ? 1 : 0;
        fprintf(trace, "%lld,%d,%.1f,%d,%d,%.4g,%d\n", q.rounds, g,
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_trace0).count(),
                q.nact, q.nbm, q.fl, dry);
      }
      if (q.nact == 0) {
        q.done = true;
        --live;
        continue;
      }
      r = run_round(h, q.c, q.nbm, true, true, sg, q.fl);
      if (r) return r;
      k_opt_step<<<q.Sg, NTHREADS, 0, sg>>>(q.c, q.a, bi, tr, lo, out);
      ++h->launches;
      r = queue_census(g, q.rounds, sg);
      if (r) return r;
      ++q.rounds;
      if (h->profiling && h->ev_used > 4000) { CK(cudaStreamSynchronize(sg)); harvest_profile(h, true, true); }
    }
    if (!progressed && !limit_hit && wait_g >= 0) CK(cudaEventSynchronize(h->gevent2[wait_g][wait_c & 1]));
  }
  if (G > 1) {
    for (int g = 0; g < G; ++g) {
      CK(cudaEventRecord(h->gevent[g], h->gstream[g]));
      CK(cudaStreamWaitEvent(st, h->gevent[g], 0));
    }
  }
  CK(cudaStreamSynchronize(st));
  if (trace) fclose(trace);
  harvest_profile(h, true, true);
  CK(cudaGetLastError());
  if (live > 0)
    return fail(GPSAT_ELIMIT, "optimiser round limit reached with " + std::to_string(live) +
                              " slot group(s) still running: unfinished experts keep status 0");
  return check_sync_timeouts(h);
}

// ------------------------------------------------------------------------------------------
// prediction
// ------------------------------------------------------------------------------------------
extern "C" int gpsat_gpr_predict(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                                 const long long* poff_host, const long long* poff_dev, const double* pcoords_dev,
                                 double* fmean_dev, double* fvar_dev, double* yvar_dev, double* fobj_dev,
                                 void* stream) {
  if (!h || !b || !theta_dev || !poff_host || !poff_dev || !pcoords_dev || !fmean_dev || !fvar_dev || !yvar_dev)
    return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int E = b->n_experts;
  long long pmax = 0;
  for (int e = 0; e < E; ++e) pmax = std::max(pmax, poff_host[e + 1] - poff_host[e]);
  const int ppmax = (int)((pmax + TB - 1) / TB) * TB;
  if (ppmax == 0) return 0;
  Plan pl;
  int r = make_plan(h, b, pl, (size_t)(MAXD + 2) * ppmax * 8 + 64);
  if (r) return r;
  Work w;
  r = setup_work(h, b, pl, w, st);
  if (r) return r;
  const int S = pl.S;
  ENS(h->pslot, (size_t)S * MAXD * ppmax * 8 + (size_t)(S + 8) * sizeof(int) + MAXD * 8);
  ENS(h->pres, (size_t)2 * S * ppmax * 8);
  // scratch for the cross-covariance tiles of one wave of items (item = slot x pair of 64-blocks)
  const size_t item_bytes = (size_t)pl.nbmax * 2 * TILE_BYTES;
  size_t scratch_budget = std::min<size_t>((size_t)8 << 30, h->budget / 8);
  int wave = (int)std::max<size_t>(1, scratch_budget / item_bytes);
  double* pslot = (double*)h->pslot.p;
  double* cs_dev = pslot + (size_t)S * MAXD * ppmax;
  int* np_dev = (int*)(cs_dev + MAXD);
  BatchIn bi = make_batch_in(b, theta_dev, (const int*)h->order.p);
  CK(cudaMemcpyAsync(cs_dev, bi.coords_scale, MAXD * 8, cudaMemcpyHostToDevice, st));
  TransformSpec tr = identity_transforms(b->D);
  std::vector<int> islot, ipb;
  for (int first = 0; first < E; first += S) {
    const int count = std::min(S, E - first);
    k_slot_init<<<S, NTHREADS, 0, st>>>(w.c, w.a, bi, tr, first, count, 0);
    ++h->launches;
    double fl = 0;
    for (int k = 0; k < count; ++k) fl += cube3(b->offsets_host, pl.order[first + k]);
    r = run_round(h, w.c, nb_of(b->offsets_host, pl.order[first]), true, false, st, fl);
    if (r) return r;
    k_pred_load<<<S, NTHREADS, 0, st>>>(w.a, count, b->D, pcoords_dev, poff_dev, cs_dev, pslot, np_dev, ppmax);
    ++h->launches;
    islot.clear();
    ipb.clear();
    for (int s = 0; s < count; ++s) {
      const int e = pl.order[first + s];
      const int npb = (int)((poff_host[e + 1] - poff_host[e] + TB - 1) / TB);
      for (int pb = 0; pb < npb; pb += 2) { islot.push_back(s); ipb.push_back(pb); }
    }
    const int n_items = (int)islot.size();
    if (n_items > 0) {
      wave = std::min(wave, n_items);
      ENS(h->items, (size_t)2 * n_items * sizeof(int));
      ENS(h->scratch, (size_t)wave * item_bytes);
      int* it_dev = (int*)h->items.p;
      // pageable copies are staged synchronously by the runtime; vectors can be reused afterwards
      CK(cudaMemcpyAsync(it_dev, islot.data(), (size_t)n_items * sizeof(int), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(it_dev + n_items, ipb.data(), (size_t)n_items * sizeof(int), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      PredCtx p;
      p.ppmax = ppmax; p.pcoords = pslot; p.np = np_dev;
      p.item_slot = it_dev; p.item_pb = it_dev + n_items;
      p.scratch = (double*)h->scratch.p;
      p.fmean = (double*)h->pres.p;
      p.fvar = (double*)h->pres.p + (size_t)S * ppmax;
      p.abuf = nullptr;
      const int nbm = nb_of(b->offsets_host, pl.order[first]);
      for (int i0 = 0; i0 < n_items; i0 += wave) {
        p.item0 = i0;
        p.n_items = std::min(wave, n_items - i0);
        k_build_xp<<<dim3(nbm, p.n_items), 256, 0, st>>>(w.c, p);
        k_predict2<false><<<p.n_items, NTHREADS, SMEM2_BYTES, st>>>(w.c, p);
        h->launches += 2;
      }
      k_pred_scatter<<<S, NTHREADS, 0, st>>>(w.c, w.a, count, poff_dev, p.fmean, p.fvar, ppmax, fmean_dev,
                                             fvar_dev, yvar_dev, fobj_dev);
      ++h->launches;
    }
  }
  CK(cudaStreamSynchronize(st));
  harvest_profile(h, true, false);
  CK(cudaGetLastError());
  return check_sync_timeouts(h);
}

// full_cov=True branch of predict (gpflow_models.py:245-263) for ONE expert (the first of the batch)
extern "C" int gpsat_gpr_predict_cov(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                                     const double* pcoords_dev, int P, double* fmean_dev, double* fcov_dev,
                                     void* stream) {
  if (!h || !b || !theta_dev || !pcoords_dev || !fmean_dev || !fcov_dev || P < 1)
    return fail(GPSAT_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  gpsat_batch b1 = *b;
  b1.n_experts = 1;
  const int ppmax = (P + TB - 1) / TB * TB;
  Plan pl;
  int r = make_plan(h, &b1, pl);
  if (r) return r;
  Work w;
  r = setup_work(h, &b1, pl, w, st);
  if (r) return r;
  const int n_items = (ppmax / TB + 1) / 2;
  const size_t item_bytes = (size_t)pl.nbmax * 2 * TILE_BYTES;
  if ((size_t)n_items * item_bytes * 2 > h->budget)
    return fail(GPSAT_ENOMEM, "full covariance at this many prediction points does not fit the memory budget");
  ENS(h->pslot, (size_t)MAXD * ppmax * 8 + 16 * sizeof(int) + MAXD * 8 + 2 * sizeof(long long));
  ENS(h->pres, (size_t)2 * ppmax * 8);
  ENS(h->scratch, (size_t)n_items * item_bytes);
  ENS(h->abuf, (size_t)n_items * item_bytes);
  ENS(h->items, (size_t)2 * n_items * sizeof(int));
  double* pslot = (double*)h->pslot.p;
  double* cs_dev = pslot + (size_t)MAXD * ppmax;
  long long* poff_dev = (long long*)(cs_dev + MAXD);
  int* np_dev = (int*)(poff_dev + 2);
  BatchIn bi = make_batch_in(&b1, theta_dev, nullptr);
  const long long poff[2] = {0, P};
  CK(cudaMemcpyAsync(cs_dev, bi.coords_scale, MAXD * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(poff_dev, poff, sizeof(poff), cudaMemcpyHostToDevice, st));
  std::vector<int> it(2 * n_items, 0);
  for (int k = 0; k < n_items; ++k) it[n_items + k] = 2 * k;
  CK(cudaMemcpyAsync(h->items.p, it.data(), it.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  TransformSpec tr = identity_transforms(b->D);
  k_slot_init<<<pl.S, NTHREADS, 0, st>>>(w.c, w.a, bi, tr, 0, 1, 0);
  r = run_round(h, w.c, pl.nbmax, true, false, st, 0.0);
  if (r) return r;
  k_pred_load<<<1, NTHREADS, 0, st>>>(w.a, 1, b->D, pcoords_dev, poff_dev, cs_dev, pslot, np_dev, ppmax);
  PredCtx p;
  p.ppmax = ppmax; p.pcoords = pslot; p.np = np_dev;
  p.item_slot = (int*)h->items.p; p.item_pb = (int*)h->items.p + n_items;
  p.n_items = n_items; p.item0 = 0;
  p.scratch = (double*)h->scratch.p;
  p.fmean = (double*)h->pres.p; p.fvar = (double*)h->pres.p + ppmax;
  p.abuf = (double*)h->abuf.p;
  k_build_xp<<<dim3(pl.nbmax, n_items), 256, 0, st>>>(w.c, p);
  k_predict2<true><<<n_items, NTHREADS, SMEM2_BYTES, st>>>(w.c, p);
  k_pred_cov<<<n_items * (n_items + 1) / 2, NTHREADS, SMEM2_BYTES, st>>>(w.c, p, fcov_dev);
  h->launches += 6;
  CK(cudaMemcpyAsync(fmean_dev, p.fmean, (size_t)P * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  harvest_profile(h, true, false);
  CK(cudaGetLastError());
  return 0;
}

#include "sgpr_host.cuh"

// ------------------------------------------------------------------------------------------
// misc entry points
// ------------------------------------------------------------------------------------------
extern "C" int gpsat_kernel_matrix(const double* x1_dev, int n1, const double* x2_dev, int n2, int D, int kernel_id,
                                   const double* theta_dev, int add_noise, double* k_dev, void* stream) {
  if (!x1_dev || !x2_dev || !theta_dev || !k_dev || D < 1 || D > MAXD) return fail(GPSAT_EINVAL, "bad argument");
  if (n1 <= 0 || n2 <= 0) return 0;
  dim3 grid((n2 + 63) / 64, (n1 + 15) / 16), block(64, 4);
  k_kernel_matrix<<<grid, block, 0, (cudaStream_t)stream>>>(x1_dev, n1, x2_dev, n2, D, kernel_id, theta_dev,
                                                            add_noise, k_dev);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_debug_factor(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev, double* l_dense,
                                  double* x_dense, void* stream) {
  if (!h || !b || !theta_dev) return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  gpsat_batch b1 = *b;
  b1.n_experts = 1;
  Plan pl;
  int r = make_plan(h, &b1, pl);
  if (r) return r;
  Work w;
  r = setup_work(h, &b1, pl, w, st);
  if (r) return r;
  BatchIn bi = make_batch_in(&b1, theta_dev, nullptr);
  TransformSpec tr = identity_transforms(b->D);
  k_slot_init<<<pl.S, NTHREADS, 0, st>>>(w.c, w.a, bi, tr, 0, 1, 0);
  // L is recycled by the inverse: dump it after a factorisation-only round, then run the full round
  if (l_dense) {
    r = run_round(h, w.c, pl.nbmax, false, false, st, 0.0);
    if (r) return r;
    k_unpack_tiles<<<dim3(pl.nbmax, pl.nbmax), 256, 0, st>>>(w.c, 0, 0, pl.nbmax, l_dense);
  }
  if (x_dense) {
    r = run_round(h, w.c, pl.nbmax, true, false, st, 0.0);
    if (r) return r;
    k_unpack_tiles<<<dim3(pl.nbmax, pl.nbmax), 256, 0, st>>>(w.c, 0, 1, pl.nbmax, x_dense);
  }
  CK(cudaStreamSynchronize(st));
  harvest_profile(h, true, false);
  CK(cudaGetLastError());
  return 0;
}

static int select_common(const gpsat_sel_spec* spec, const double* table_dev, long long n, const double* refs_dev,
                         int nrefcols, int E, int fill, long long* counts, const long long* offsets, int* idx,
                         void* stream) {
  if (!spec || !table_dev || !refs_dev || nrefcols < 1 || nrefcols > 16 || spec->nterms < 0 ||
      spec->nterms > SEL_MAXTERMS)
    return fail(GPSAT_EINVAL, "bad argument");
  if (E <= 0) return 0;
  SelSpec sp;
  memcpy(&sp, spec, sizeof(sp));
  k_select<<<E, 256, 0, (cudaStream_t)stream>>>(sp, table_dev, (long)n, refs_dev, nrefcols, fill, counts, offsets,
                                                idx);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int gpsat_select_count(const gpsat_sel_spec* spec, const double* table_dev, long long n,
                                  const double* refs_dev, int nrefcols, int E, long long* counts_dev, void* stream) {
  if (!counts_dev) return fail(GPSAT_EINVAL, "counts_dev is NULL");
  return select_common(spec, table_dev, n, refs_dev, nrefcols, E, 0, counts_dev, nullptr, nullptr, stream);
}
extern "C" int gpsat_select_fill(const gpsat_sel_spec* spec, const double* table_dev, long long n,
                                 const double* refs_dev, int nrefcols, int E, const long long* offsets_dev,
                                 int* idx_dev, void* stream) {
  if (!offsets_dev || !idx_dev) return fail(GPSAT_EINVAL, "null output");
  return select_common(spec, table_dev, n, refs_dev, nrefcols, E, 1, nullptr, offsets_dev, idx_dev, stream);
}

// ---- grid-bucketed selection ----
static int find_ball_term(const gpsat_sel_spec* spec) {
  for (int k = 0; k < spec->nterms; ++k)
    if ((spec->t[k].type == 1 || spec->t[k].type == 2) && spec->t[k].ncol == 2) return k;
  return -1;
}
static CellGrid make_grid(const gpsat_sel_spec* spec, int term, const gpsat_cell_grid* cg) {
  CellGrid g;
  g.x0 = cg->x0; g.y0 = cg->y0; g.inv_cell = 1.0 / cg->cell; g.ncx = cg->ncx; g.ncy = cg->ncy;
  g.colx = spec->t[term].col[0]; g.coly = spec->t[term].col[1];
  g.rcolx = spec->t[term].rcol[0]; g.rcoly = spec->t[term].rcol[1];
  return g;
}

extern "C" int gpsat_bucket_build(const gpsat_sel_spec* spec, const gpsat_cell_grid* cg, const double* table_dev,
                                  long long n, int pass, int* counts_dev, const long long* start_dev,
                                  int* order_dev, void* stream) {
  if (!spec || !cg || !table_dev || !counts_dev) return fail(GPSAT_EINVAL, "null argument");
  const int term = find_ball_term(spec);
  if (term < 0) return fail(GPSAT_EINVAL, "no two-column ball / max_dist term to bucket on");
  if (!(cg->cell >= spec->t[term].val * 1.0000001)) return fail(GPSAT_EINVAL, "cell edge must exceed the radius");
  if (n <= 0) return 0;
  CellGrid g = make_grid(spec, term, cg);
  const int grid = (int)std::min<long long>((n + 255) / 256, 148LL * 8);
  if (pass == 0) {
    k_cell_hist<<<grid, 256, 0, (cudaStream_t)stream>>>(g, table_dev, (long)n, counts_dev);
  } else {
    if (!start_dev || !order_dev) return fail(GPSAT_EINVAL, "null argument");
    k_cell_scatter<<<grid, 256, 0, (cudaStream_t)stream>>>(g, table_dev, (long)n, start_dev, counts_dev, order_dev);
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_select_bucket(const gpsat_sel_spec* spec, const gpsat_cell_grid* cg, const double* table_dev,
                                   long long n, const double* refs_dev, int nrefcols, int n_experts,
                                   const long long* start_dev, const int* order_dev, int max_count,
                                   long long* counts_dev, const long long* offsets_dev, int* idx_dev, void* stream) {
  if (!spec || !cg || !table_dev || !refs_dev || !start_dev || !order_dev || nrefcols < 1 || nrefcols > 16)
    return fail(GPSAT_EINVAL, "bad argument");
  const int term = find_ball_term(spec);
  if (term < 0) return fail(GPSAT_EINVAL, "no two-column ball / max_dist term to bucket on");
  if (n_experts <= 0) return 0;
  const int fill = counts_dev == nullptr;
  if (fill && (!offsets_dev || !idx_dev)) return fail(GPSAT_EINVAL, "null output");
  int cap = 1;
  while (cap < std::max(max_count, 1)) cap <<= 1;
  if (fill && cap > 32768) return fail(GPSAT_EINVAL, "more than 32768 matches per expert: use gpsat_select_fill");
  SelSpec sp;
  memcpy(&sp, spec, sizeof(sp));
  CellGrid g = make_grid(spec, term, cg);
  const size_t smem = fill ? (size_t)cap * sizeof(int) : 0;
  // (the dynamic shared-memory attribute of k_select_bucket is set per device in gpsat_create; a caller that
  //  uses the selection entry points without a handle on this device gets it set here)
  {
    int dev = 0;
    static bool attr_dev[64] = {};
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_dev[dev]) {
      CK(cudaFuncSetAttribute(k_select_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * sizeof(int)));
      attr_dev[dev] = true;
    }
  }
  k_select_bucket<<<n_experts, 256, smem, (cudaStream_t)stream>>>(sp, g, table_dev, (long)n, refs_dev, nrefcols,
                                                                  start_dev, order_dev, fill, cap, counts_dev,
                                                                  offsets_dev, idx_dev);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_gather_rows(const double* table_dev, long long n, const int* idx_dev, long long total,
                                 int D, const int* coord_cols, int obs_col, double* coords_dev, double* obs_dev,
                                 void* stream) {
  if (!table_dev || !coords_dev || !coord_cols || D < 1 || D > MAXD) return fail(GPSAT_EINVAL, "bad argument");
  if (total <= 0) return 0;
  if (!idx_dev) return fail(GPSAT_EINVAL, "idx_dev is NULL");
  GatherCols gc;
  gc.D = D;
  gc.ocol = obs_col;
  for (int d = 0; d < MAXD; ++d) gc.ccols[d] = d < D ? coord_cols[d] : 0;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
  k_gather_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(gc, table_dev, (long)n, idx_dev, (long)total, coords_dev,
                                                        obs_col >= 0 ? obs_dev : nullptr);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_gather_pred(const double* table_dev, long long n, const double* refs_dev, int nrefcols,
                                 int n_experts, const long long* offsets_dev, const int* idx_dev, int D,
                                 const int* table_cols, const int* ref_cols, double* out_dev, void* stream) {
  if (!table_dev || !refs_dev || !offsets_dev || !out_dev || !table_cols || !ref_cols || D < 1 || D > MAXD)
    return fail(GPSAT_EINVAL, "bad argument");
  if (n_experts <= 0) return 0;
  PredCols pc;
  pc.D = D;
  for (int d = 0; d < MAXD; ++d) {
    pc.tcol[d] = d < D ? table_cols[d] : -1;
    pc.rcol[d] = d < D ? ref_cols[d] : 0;
  }
  k_gather_pred<<<n_experts, 256, 0, (cudaStream_t)stream>>>(pc, table_dev, (long)n, refs_dev, nrefcols,
                                                             offsets_dev, idx_dev, out_dev);
  CK(cudaGetLastError());
  return 0;
}

// ---- FP64 tensor-pipe speed of light: register-resident DMMA chains, no memory traffic ----
__global__ void __launch_bounds__(256) k_dmma_peak(int iters, double* sink) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) sink[0] = s;
}
extern "C" int gpsat_dmma_peak(int device, int iters, double* tflops_out, double* ms_out) {
  if (!tflops_out || iters < 1) return fail(GPSAT_EINVAL, "bad argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double* sink = nullptr;
  CK(cudaMalloc(&sink, 8));
  const int grid = prop.multiProcessorCount * 4;   // 4 x 8 warps per SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_dmma_peak<<<grid, 256>>>(iters / 8 + 1, sink);  // warm-up
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dmma_peak<<<grid, 256>>>(iters, sink);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, (double)ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const double flops = 2.0 * 8 * 8 * 4 * 8.0 * (double)iters * (256 / 32) * (double)grid;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  CK(cudaGetLastError());
  return 0;
}

// ---- micro-benchmarks (see microbench.cuh): which = 0..3 chains(1,2,4,8 accumulators) with `param` CTAs
// per SM; 10/11 = 64x64 / 128x128 core, param = mode (0 smem, 1 HBM stream, 2 L2 shared), nk k-tiles ----
template <class F>
static int time_launch(F launch, double* ms_out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, (double)ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = best;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_microbench(int device, int which, int param, int nk, double* tflops_out) {
  if (!tflops_out) return fail(GPSAT_EINVAL, "bad argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int nsm = prop.multiProcessorCount;
  double* buf = nullptr;
  double ms = 0;
  int r = 0;
  if (which < 4) {
    CK(cudaMalloc(&buf, 64));
    const int grid = nsm * std::max(1, param), iters = nk;
    const int nch = 1 << which;
    if (which == 0) r = time_launch([&] { k_dmma_chain<1><<<grid, 256>>>(iters, buf); }, &ms);
    if (which == 1) r = time_launch([&] { k_dmma_chain<2><<<grid, 256>>>(iters, buf); }, &ms);
    if (which == 2) r = time_launch([&] { k_dmma_chain<4><<<grid, 256>>>(iters, buf); }, &ms);
    if (which == 3) r = time_launch([&] { k_dmma_chain<8><<<grid, 256>>>(iters, buf); }, &ms);
    *tflops_out = 2.0 * 256 * nch * (double)iters * 8 * grid / (ms * 1e-3) / 1e12;
  } else if (which == 5 || which == 6) {   // DMMA chains from ONE warp per scheduler (128-thread CTAs, one per SM)
    CK(cudaMalloc(&buf, 64));
    const int iters = nk;
    if (which == 5) r = time_launch([&] { k_dmma_chain<8><<<nsm, 128>>>(iters, buf); }, &ms);
    else r = time_launch([&] { k_dmma_chain<32><<<nsm, 128>>>(iters, buf); }, &ms);
    *tflops_out = 2.0 * 256 * (which == 5 ? 8 : 32) * (double)iters * 4 * nsm / (ms * 1e-3) / 1e12;
  } else if (which == 50 || which == 51) {
    // task streams of nk k-tiles each (param = tasks per CTA) with the 128x128 core (51)
    const long tiles_per_cta = 48;
    const int ctas = nsm;
    const size_t bytes = (size_t)ctas * tiles_per_cta * TILE_BYTES;
    CK(cudaMalloc(&buf, bytes + (size_t)ctas * 4 * TILE_BYTES));
    CK(cudaMemset(buf, 0, bytes));
    const int tasks = std::max(1, param);
    if (which == 50) {
      cudaFree(buf);
      return fail(GPSAT_EINVAL, "the 128x64 two-CTA core was an experiment of round 1 and is no longer built");    } else {
      auto kern = k_gemm2_tasks_bench<false, false>;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_ELEMS * 8));
      r = time_launch([&] { kern<<<ctas, NTHREADS, G2_SMEM_ELEMS * 8>>>(buf, tiles_per_cta, nk, tasks, buf + bytes / 8); }, &ms);
      *tflops_out = 2.0 * 128 * 128 * 64 * (double)nk * tasks * ctas / (ms * 1e-3) / 1e12;
    }
  } else if (which == 40) {   // max relative error of exp_neg (param 0) / sqrt_pos (param 1) vs libdevice
    CK(cudaMalloc(&buf, 64));
    CK(cudaMemset(buf, 0, 64));
    k_elem_accuracy<<<nsm * 8, 256>>>(param, std::max(1, nk), (unsigned long long*)buf);
    CK(cudaGetLastError());
    CK(cudaMemcpy(tflops_out, buf, 8, cudaMemcpyDeviceToHost));
  } else if (which == 30 || which == 31) {
    // 30: microseconds per 128x128 diagonal block (one CTA per SM, nk repetitions)
    // 31: SM cycles from the entry of diag_block_128 to its stage `param` (1..9, see ClockProbe), last repetition
    CK(cudaMalloc(&buf, (size_t)nsm * (6 * TILE_BYTES + 64) + 256));
    CK(cudaMemset(buf, 0, (size_t)nsm * (6 * TILE_BYTES + 64) + 256));
    double* ld = buf + (size_t)nsm * 6 * TILE_ELEMS;
    int* fl = (int*)(ld + 2 * nsm);
    long long* stamps = (long long*)(buf + (size_t)nsm * (6 * TILE_ELEMS + 8));
    CK(cudaFuncSetAttribute(k_diag_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BYTES));
    r = time_launch([&] {
      k_diag_bench<<<nsm, NTHREADS, PANEL_SMEM_BYTES>>>(nk, buf, fl, ld, which == 31 ? stamps : nullptr);
    }, &ms);
    if (which == 30) {
      *tflops_out = ms * 1e3 / nk;
    } else {
      long long st[16];
      CK(cudaMemcpy(st, stamps, sizeof(st), cudaMemcpyDeviceToHost));
      *tflops_out = (double)(st[std::max(0, std::min(15, param))] - st[0]);
    }
  } else if (which == 20 || which == 21) {
    CK(cudaMalloc(&buf, 64));
    const int grid = nsm * 4;
    if (which == 20) {
      r = time_launch([&] { k_pipe_mix<<<grid, 256>>>(nk, param, buf); }, &ms);
      *tflops_out = ms;   // milliseconds: compare modes 1, 2, 3
    } else {
      r = time_launch([&] { k_kern_rate<<<grid, 256>>>(nk, param, buf); }, &ms);
      *tflops_out = (double)grid * 256 * nk / (ms * 1e-3) / 1e9;   // G entries / s
    }
  } else {
    const long tiles_per_cta = 64;
    const size_t bytes = (size_t)(param == 1 ? nsm : 1) * tiles_per_cta * TILE_BYTES;
    CK(cudaMalloc(&buf, bytes + 64));
    CK(cudaMemset(buf, 0, bytes + 64));
    if (which == 10) {
      CK(cudaFuncSetAttribute(k_gemm1_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TILE_BYTES));
      r = time_launch([&] { k_gemm1_bench<<<nsm, NTHREADS, 4 * TILE_BYTES>>>(buf, tiles_per_cta, nk, param, buf + bytes / 8); }, &ms);
      *tflops_out = 2.0 * 64 * 64 * 64 * (double)nk * nsm / (ms * 1e-3) / 1e12;
    } else {
      auto kern = (which == 11) ? k_gemm2_bench<false, false> : (which == 12 ? k_gemm2_bench<true, true> : k_gemm2_bench<false, true>);
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_ELEMS * 8));
      r = time_launch([&] { kern<<<nsm, NTHREADS, G2_SMEM_ELEMS * 8>>>(buf, tiles_per_cta, nk, param, buf + bytes / 8); }, &ms);
      *tflops_out = 2.0 * 128 * 128 * 64 * (double)nk * nsm / (ms * 1e-3) / 1e12;
    }
  }
  cudaFree(buf);
  return r;
}

// ---- post-processing ----
extern "C" int gpsat_gaussian_smooth(const double* qx_dev, const double* qy_dev, const int* seg_of_query_dev,
                                     long long n_query, const double* x_dev, const double* y_dev,
                                     const double* vals_dev, const long long* seg_off_dev, double l_x, double l_y,
                                     const double* vmin, const double* vmax, double* out_dev, void* stream) {
  if (!qx_dev || !qy_dev || !seg_of_query_dev || !x_dev || !y_dev || !vals_dev || !seg_off_dev || !out_dev ||
      !(l_x > 0) || !(l_y > 0))
    return fail(GPSAT_EINVAL, "bad argument");
  if (n_query <= 0) return 0;
  k_gauss_smooth<<<(unsigned)n_query, 256, 0, (cudaStream_t)stream>>>(
      qx_dev, qy_dev, seg_of_query_dev, x_dev, y_dev, vals_dev, seg_off_dev, l_x, l_y, vmin ? *vmin : 0.0,
      vmax ? *vmax : 0.0, vmin != nullptr, vmax != nullptr, out_dev);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_weighted_groups(const double* ref_dev, const double* to_dev, int nd, const double* vals_dev,
                                     long long n, int ncol, const long long* order_dev,
                                     const long long* group_off_dev, long long n_groups, double lengthscale,
                                     double* out_dev, void* stream) {
  if (!ref_dev || !to_dev || nd < 1 || (ncol > 0 && !vals_dev) || ncol < 0 || !order_dev || !group_off_dev ||
      !out_dev || !(lengthscale > 0))
    return fail(GPSAT_EINVAL, "bad argument");
  if (n_groups <= 0) return 0;
  k_weighted_groups<<<(unsigned)n_groups, 256, 0, (cudaStream_t)stream>>>(
      ref_dev, to_dev, nd, vals_dev, (long)n, ncol, order_dev, group_off_dev, lengthscale * lengthscale,
      (long)n_groups, out_dev);
  CK(cudaGetLastError());
  return 0;
}

// ---- upstream binning ----
extern "C" int gpsat_bin_accumulate(const double* x_dev, const double* y_dev, const double* vals_dev,
                                    const int* group_dev, long long n, const double* x_edges_dev, int n_x_edges,
                                    double x_round_scale, int x_round_div, const double* y_edges_dev, int n_y_edges,
                                    double y_round_scale, int y_round_div, int n_groups, double* sum_dev,
                                    unsigned long long* count_dev, void* stream) {
  if (!x_dev || !vals_dev || !x_edges_dev || n_x_edges < 2 || !sum_dev || !count_dev || n_groups < 1 ||
      (y_dev && (!y_edges_dev || n_y_edges < 2)))
    return fail(GPSAT_EINVAL, "bad argument");
  if (n <= 0) return 0;
  BinAxis ax{x_edges_dev, n_x_edges, x_round_scale, x_round_div};
  BinAxis ay{y_edges_dev, n_y_edges, y_round_scale, y_round_div};
  const long long want = (n + 255) / 256;
  const unsigned grid = (unsigned)std::min<long long>(want, 148LL * 16);
  k_bin_accumulate<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, vals_dev, group_dev, n, ax, ay, y_dev != nullptr,
                                                         sum_dev, count_dev);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_bin_spread(const double* x_dev, const double* y_dev, const double* vals_dev, const int* group_dev,
                                long long n, const double* x_edges_dev, int n_x_edges, double x_round_scale,
                                int x_round_div, const double* y_edges_dev, int n_y_edges, double y_round_scale,
                                int y_round_div, int n_groups, const double* sum_dev,
                                const unsigned long long* count_dev, double* ssd_dev, double* min_dev, double* max_dev,
                                void* stream) {
  if (!x_dev || !vals_dev || !x_edges_dev || n_x_edges < 2 || n_groups < 1 || (y_dev && (!y_edges_dev || n_y_edges < 2)) ||
      (ssd_dev && (!sum_dev || !count_dev)) || (!ssd_dev && !min_dev && !max_dev))
    return fail(GPSAT_EINVAL, "bad argument");
  if (n <= 0) return 0;
  BinAxis ax{x_edges_dev, n_x_edges, x_round_scale, x_round_div};
  BinAxis ay{y_edges_dev, n_y_edges, y_round_scale, y_round_div};
  const long long want = (n + 255) / 256;
  const unsigned grid = (unsigned)std::min<long long>(want, 148LL * 16);
  k_bin_spread<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, vals_dev, group_dev, n, ax, ay, y_dev != nullptr,
                                                     sum_dev, count_dev, ssd_dev, min_dev, max_dev);
  CK(cudaGetLastError());
  return 0;
}

// ---- host-side L-BFGS hooks (same code the device runs) ----
extern "C" size_t gpsat_lbfgs_state_bytes(void) { return sizeof(LbfgsState); }
extern "C" void gpsat_lbfgs_init_host(void* state, const double* x0, int n) {
  lb_init(*reinterpret_cast<LbfgsState*>(state), x0, n);
}
extern "C" int gpsat_lbfgs_tell_host(void* state, const gpsat_opt_options* o, double f, const double* g,
                                     double* x_next, int* nit, int* nfev) {
  LbfgsState& st = *reinterpret_cast<LbfgsState*>(state);
  LbfgsOpts lo;
  lo.m = o->maxcor; lo.maxiter = o->maxiter; lo.maxfun = o->maxfun; lo.maxls = o->maxls;
  lo.factr = o->ftol / 2.220446049250313e-16;
  lo.pgtol = o->gtol;
  lbfgs_tell(st, lo, f, g);
  for (int i = 0; i < st.n; ++i) x_next[i] = st.x[i];
  if (nit) *nit = st.nit;
  if (nfev) *nfev = st.nfev;
  return st.status;
}
