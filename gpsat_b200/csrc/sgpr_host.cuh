// gpsat_b200: host orchestration of the batched sparse GPR (included by api.cu; see sgpr.cuh).
#pragma once

struct SgWork {
  Plan pl;
  Work w;        // cz: the Z-slot pool (Kuu) -- also the optimiser's slot pool
  SlotCtx cb;    // the B pool (shares n / nb / active / theta with cz)
  SgCtx g;
  gpsat_batch zb;   // the inducing points seen as an exact-GPR batch (coords = Z, obs = 0)
  int nbM, mbmax, ncolmax;
};

__global__ void k_sg_ymean(const double* __restrict__ y, const long long* __restrict__ off, int E, int local,
                           double* __restrict__ out) {
  const int e = blockIdx.x;
  __shared__ double red[NTHREADS / 32];
  double v[1] = {0.0};
  if (local) {
    for (long long i = off[e] + threadIdx.x; i < off[e + 1]; i += NTHREADS) v[0] += y[i];
    block_sum<1>(v, red);
  }
  if (threadIdx.x == 0) out[e] = local ? v[0] / (double)(off[e + 1] - off[e]) : 0.0;
}

static int sg_setup(gpsat_handle* h, const gpsat_sgpr_batch* sb, long long pmax, SgWork& W, cudaStream_t st) {
  const gpsat_batch* b = &sb->data;
  const int E = b->n_experts;
  if (E <= 0) return fail(GPSAT_EINVAL, "n_experts must be positive");
  if (!sb->z_offsets_host || !sb->z_offsets_dev || !sb->z_coords_dev) return fail(GPSAT_EINVAL, "null inducing points");
  long long nmax = 0, mmax = 0, summ = sb->z_offsets_host[E];
  for (int e = 0; e < E; ++e) {
    nmax = std::max(nmax, b->offsets_host[e + 1] - b->offsets_host[e]);
    mmax = std::max(mmax, sb->z_offsets_host[e + 1] - sb->z_offsets_host[e]);
  }
  ENS(h->sg_zeros, (size_t)std::max<long long>(summ, 1) * 8);
  CK(cudaMemsetAsync(h->sg_zeros.p, 0, (size_t)std::max<long long>(summ, 1) * 8, st));
  W.zb = *b;
  W.zb.offsets_host = sb->z_offsets_host;
  W.zb.offsets_dev = sb->z_offsets_dev;
  W.zb.coords_dev = sb->z_coords_dev;
  W.zb.obs_dev = (const double*)h->sg_zeros.p;
  W.zb.obs_mean_local = 0;
  W.zb.obs_mean_out_dev = nullptr;
  W.mbmax = (int)((mmax + TB - 1) / TB);
  const int mrows = W.mbmax + 1;                                   // room for the augmented row
  W.ncolmax = (int)((std::max(nmax, pmax) + TB - 1) / TB);
  const size_t mm_tiles = (size_t)mrows * mrows, mn_tiles = (size_t)mrows * W.ncolmax;
  const int vlen = (std::max(W.ncolmax, mrows) + 1) * TB;
  const int gp_n = W.mbmax * W.ncolmax + W.mbmax * W.mbmax;
  const size_t extra = (4 * mm_tiles + 2 * mn_tiles) * TILE_BYTES + (size_t)8 * vlen * 8 + (size_t)gp_n * NG * 8 +
                       3 * ((size_t)(mrows + 1) * (mrows + 2) / 2) * TILE_BYTES + 4096;
  int r = make_plan(h, &W.zb, W.pl, extra);
  if (r) return r;
  r = setup_work(h, &W.zb, W.pl, W.w, st);
  if (r) return r;
  const int S = W.pl.S;
  W.nbM = W.pl.nbmax;
  SlotCtx& cz = W.w.c;
  cz.nvar_override = SGPR_JITTER;
  ENS(h->Lt2, (size_t)S * W.pl.ntmax * TILE_BYTES);
  ENS(h->Xt2, (size_t)S * W.pl.ntmax * TILE_BYTES);
  ENS(h->Kt2, (size_t)S * W.pl.ntmax * TILE_BYTES);
  ENS(h->quad2, (size_t)S * 8);
  ENS(h->logdet2, (size_t)S * W.pl.nbmax * 8);
  ENS(h->fail2, (size_t)2 * S * sizeof(int));
  CK(cudaMemsetAsync(h->fail2.p, 0, (size_t)2 * S * sizeof(int), st));
  W.cb = cz;
  W.cb.nvar_override = -1.0;
  W.cb.Lt = (double*)h->Lt2.p; W.cb.Xt = (double*)h->Xt2.p; W.cb.Kt = (double*)h->Kt2.p;
  W.cb.quad = (double*)h->quad2.p; W.cb.logdet_part = (double*)h->logdet2.p; W.cb.fail = (int*)h->fail2.p; W.cb.pflag = (int*)h->fail2.p + S;
  ENS(h->sg_mm, (size_t)S * 4 * mm_tiles * TILE_BYTES);
  ENS(h->sg_mn, (size_t)S * 2 * mn_tiles * TILE_BYTES);
  ENS(h->sg_vec, (size_t)S * 8 * vlen * 8);
  ENS(h->sg_scal, (size_t)S * 16 * 8);
  ENS(h->sg_gpart, (size_t)S * gp_n * NG * 8);
  ENS(h->sg_ints, (size_t)4 * S * sizeof(int));
  ENS(h->sg_beta, (size_t)S * 8);
  ENS(h->sg_ymean, (size_t)E * 8);
  SgCtx& g = W.g;
  g.S = S; g.D = b->D; g.kid = b->kernel_id; g.mbmax = W.mbmax; g.nbnmax = W.ncolmax;
  g.slot_expert = W.w.a.slot_expert; g.active = cz.active; g.theta = cz.theta;
  g.zcoords = cz.coords; g.npmax_z = cz.npmax; g.mz = cz.n;
  g.xcoords = b->coords_dev; g.yobs = b->obs_dev; g.xoff = b->offsets_dev;
  for (int d = 0; d < MAXD; ++d) g.coords_scale[d] = (d < b->D && b->coords_scale[d] != 0.0) ? b->coords_scale[d] : 1.0;
  g.obs_scale = (b->obs_scale != 0.0) ? b->obs_scale : 1.0;
  g.ymean = (const double*)h->sg_ymean.p;
  double* mm = (double*)h->sg_mm.p;
  const long mm_stride = (long)mm_tiles * TILE_ELEMS;
  g.XLF = TMat{mm, 4 * mm_stride, mrows};
  g.BF = TMat{mm + mm_stride, 4 * mm_stride, mrows};
  g.EF = TMat{mm + 2 * mm_stride, 4 * mm_stride, mrows};
  g.WF = TMat{mm + 3 * mm_stride, 4 * mm_stride, mrows};
  double* mn = (double*)h->sg_mn.p;
  const long mn_stride = (long)mn_tiles * TILE_ELEMS;
  g.KUF = TMat{mn, 2 * mn_stride, W.ncolmax};
  g.AP = TMat{mn + mn_stride, 2 * mn_stride, W.ncolmax};
  g.vec = (double*)h->sg_vec.p; g.vlen = vlen;
  g.scal = (double*)h->sg_scal.p;
  g.gpart = (double*)h->sg_gpart.p; g.gp_n = gp_n;
  int* ip = (int*)h->sg_ints.p;
  g.mb = ip; g.mb1 = ip + S; g.nbn = ip + 2 * S; g.npb = ip + 3 * S;
  g.beta = (double*)h->sg_beta.p;
  k_sg_ymean<<<E, NTHREADS, 0, st>>>(b->obs_dev, b->offsets_dev, E, b->obs_mean_local, (double*)h->sg_ymean.p);
  ++h->launches;
  if (b->obs_mean_out_dev)
    CK(cudaMemcpyAsync(b->obs_mean_out_dev, h->sg_ymean.p, (size_t)E * 8, cudaMemcpyDeviceToDevice, st));
  return 0;
}

template <bool TA, bool TBm>
static void sg_gemm(gpsat_handle* h, const SgWork& W, TMat A, TMat B, TMat C, const int* mt, const int* nt,
                    const int* kt, int mt_max, int nt_max, int kmode, int flags, cudaStream_t st) {
  TGemm p;
  p.A = A; p.B = B; p.C = C; p.mt = mt; p.nt = nt; p.kt = kt; p.active = W.g.active; p.beta = W.g.beta;
  p.kmode = kmode; p.flags = flags; p.diag_limit_from_mz = 0; p.mz = W.g.mz;
  const int nis = (mt_max + 1) / 2, njs = (nt_max + 1) / 2;
  k_tgemm<TA, TBm><<<dim3(nis * njs, W.g.S), NTHREADS, SMEM2_BYTES, st>>>(p, njs);
  ++h->launches;
}

// one evaluation of -ELBO (and its gradient) for all active slots; leaves L, X_L in cz and LB (X_B) in cb
static int sg_round(gpsat_handle* h, SgWork& W, bool grad, bool inverse_b, cudaStream_t st) {
  const SlotCtx& cz = W.w.c;
  const SgCtx& g = W.g;
  const int S = g.S, mb = W.mbmax, mr = W.mbmax + 1, nc = W.ncolmax;
  int r = run_round_flags(h, cz, W.nbM, RR_BUILD | RR_INVERSE, st, 0.0);
  if (r) return r;
  k_sg_setup<<<S, 256, 0, st>>>(g);
  k_sg_unpack_x<<<dim3(mr, mr, S), 256, 0, st>>>(cz, g, g.XLF, 0);
  k_sg_build_uf<<<dim3(nc, mb, S), 256, 0, st>>>(g, g.KUF, nullptr, nullptr);
  h->launches += 3;
  sg_gemm<false, true>(h, W, g.XLF, g.KUF, g.AP, g.mb, g.nbn, g.mb, mb, nc, 1, 0, st);          // A' = X_L Kuf
  sg_gemm<false, false>(h, W, g.AP, g.AP, g.BF, g.mb, g.mb, g.nbn, mb, mb, 0, 1 | 2 | 4, st);    // B = I + beta A'A'^T
  k_sg_ay<<<dim3(mb, S), 256, 0, st>>>(g);
  k_sg_pack_b<<<dim3(W.pl.ntmax, S), 256, 0, st>>>(W.cb, g);
  h->launches += 2;
  r = run_round_flags(h, W.cb, W.nbM, ((grad || inverse_b) ? RR_INVERSE : 0) | (grad ? RR_LAUUM : 0), st, 0.0);
  if (r) return r;
  if (grad) {
    k_sg_prep<<<dim3(mb, mb, S), 256, 0, st>>>(W.cb, g);
    k_sg_vec2<<<dim3(mb + nc, S), 256, 0, st>>>(g);
    h->launches += 2;
    sg_gemm<true, true>(h, W, g.XLF, g.EF, g.BF, g.mb, g.mb, g.mb, mb, mb, 2, 0, st);             // T = X_L' E
    sg_gemm<false, true>(h, W, g.BF, g.AP, g.KUF, g.mb, g.nbn, g.mb, mb, nc, 0, 0, st);           // GUF = T A'
    sg_gemm<false, true>(h, W, g.WF, g.XLF, g.EF, g.mb, g.mb, g.mb, mb, mb, 3, 0, st);            // U = W X_L
    sg_gemm<true, true>(h, W, g.XLF, g.EF, g.WF, g.mb, g.mb, g.mb, mb, mb, 2, 0, st);             // G1 = X_L' U
    k_sg_trace<false><<<dim3(nc, mb, S), 256, 0, st>>>(g, g.KUF);
    k_sg_trace<true><<<dim3(mb, mb, S), 256, 0, st>>>(g, g.WF);
    h->launches += 2;
  }
  k_sg_finalize<<<S, 256, 0, st>>>(cz, W.cb, g, grad ? 1 : 0);
  ++h->launches;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gpsat_sgpr_eval(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta_dev, double* f_dev,
                               double* grad_dev, void* stream) {
  if (!h || !sb || !theta_dev) return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  SgWork W;
  int r = sg_setup(h, sb, 0, W, st);
  if (r) return r;
  const gpsat_batch* b = &sb->data;
  BatchIn bi = make_batch_in(&W.zb, theta_dev, (const int*)h->order.p);
  TransformSpec tr = identity_transforms(b->D);
  for (int first = 0; first < b->n_experts; first += W.pl.S) {
    const int count = std::min(W.pl.S, b->n_experts - first);
    k_slot_init<<<W.pl.S, NTHREADS, 0, st>>>(W.w.c, W.w.a, bi, tr, first, count, 0);
    ++h->launches;
    r = sg_round(h, W, grad_dev != nullptr, false, st);
    if (r) return r;
    k_eval_scatter<<<(count + 127) / 128, 128, 0, st>>>(W.w.c, W.w.a, count, f_dev, grad_dev, b->D + 2);
    ++h->launches;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  return check_sync_timeouts(h);
}

extern "C" int gpsat_sgpr_optimise(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta0_dev,
                                   const gpsat_transforms* trs, const gpsat_opt_options* opts, double* theta_out_dev,
                                   double* fobj_out_dev, int* status_out_dev, int* nit_out_dev, int* nfev_out_dev,
                                   void* stream) {
  if (!h || !sb || !theta0_dev || !trs || !theta_out_dev || !fobj_out_dev || !status_out_dev || !nit_out_dev ||
      !nfev_out_dev)
    return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  gpsat_opt_options od;
  gpsat_default_opts(&od);
  if (opts) od = *opts;
  if (od.maxcor < 1 || od.maxcor > LB_M) return fail(GPSAT_EINVAL, "maxcor must be in [1, 10]");
  SgWork W;
  int r = sg_setup(h, sb, 0, W, st);
  if (r) return r;
  const gpsat_batch* b = &sb->data;
  BatchIn bi = make_batch_in(&W.zb, theta0_dev, (const int*)h->order.p);
  TransformSpec tr;
  memset(&tr, 0, sizeof(tr));
  tr.np = b->D + 2;
  for (int p = 0; p < tr.np; ++p) {
    tr.kind[p] = trs->kind[p];
    tr.low[p] = trs->low[p];
    tr.high[p] = trs->high[p];
    if (trs->trainable[p]) tr.free_idx[tr.nfree++] = p;
  }
  if (tr.nfree == 0) return fail(GPSAT_EINVAL, "no trainable parameters (use gpsat_sgpr_eval)");
  LbfgsOpts lo;
  lo.m = od.maxcor; lo.maxiter = od.maxiter; lo.maxfun = od.maxfun; lo.maxls = od.maxls;
  lo.factr = od.ftol / 2.220446049250313e-16;
  lo.pgtol = od.gtol;
  OptOut out{theta_out_dev, fobj_out_dev, status_out_dev, nit_out_dev, nfev_out_dev};
  const int S = W.pl.S;
  const int count = std::min(S, b->n_experts);
  k_slot_init<<<S, NTHREADS, 0, st>>>(W.w.c, W.w.a, bi, tr, 0, count, 1);
  ++h->launches;
  CK(cudaMemcpyAsync(W.w.a.queue_head, &count, sizeof(int), cudaMemcpyHostToDevice, st));
  const long long max_rounds = (long long)(b->n_experts / S + 2) * (od.maxfun + od.maxls + 2);
  bool finished = false;
  for (long long round = 0; round < max_rounds; ++round) {
    CK(cudaMemcpyAsync(h->host_ints, W.w.c.n, (size_t)3 * S * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int nact = 0;
    for (int s = 0; s < S; ++s) nact += h->host_ints[2 * S + s] ? 1 : 0;
    if (nact == 0) { finished = true; break; }
    r = sg_round(h, W, true, false, st);
    if (r) return r;
    k_opt_step<<<S, NTHREADS, 0, st>>>(W.w.c, W.w.a, bi, tr, lo, out);
    ++h->launches;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (!finished)
    return fail(GPSAT_ELIMIT, "optimiser round limit reached with experts still running (status 0)");
  return check_sync_timeouts(h);
}

// grid (S): ceil(P/64) per slot
__global__ void k_sg_npb(SgCtx g, const long long* poff) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.S || !g.active[s]) return;
  const int e = g.slot_expert[s];
  g.npb[s] = (int)((poff[e + 1] - poff[e] + TB - 1) / TB);
}

extern "C" int gpsat_sgpr_predict(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta_dev,
                                  const long long* poff_host, const long long* poff_dev, const double* pcoords_dev,
                                  double* fmean_dev, double* fvar_dev, double* yvar_dev, double* fobj_dev,
                                  void* stream) {
  if (!h || !sb || !theta_dev || !poff_host || !poff_dev || !pcoords_dev || !fmean_dev || !fvar_dev || !yvar_dev)
    return fail(GPSAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const gpsat_batch* b = &sb->data;
  const int E = b->n_experts;
  long long pmax = 0;
  for (int e = 0; e < E; ++e) pmax = std::max(pmax, poff_host[e + 1] - poff_host[e]);
  if (pmax == 0) return 0;
  SgWork W;
  int r = sg_setup(h, sb, pmax, W, st);
  if (r) return r;
  const SgCtx& g = W.g;
  BatchIn bi = make_batch_in(&W.zb, theta_dev, (const int*)h->order.p);
  TransformSpec tr = identity_transforms(b->D);
  const int S = W.pl.S, mb = W.mbmax, mr = mb + 1, nc = W.ncolmax;
  for (int first = 0; first < E; first += S) {
    const int count = std::min(S, E - first);
    k_slot_init<<<S, NTHREADS, 0, st>>>(W.w.c, W.w.a, bi, tr, first, count, 0);
    ++h->launches;
    r = sg_round(h, W, false, true, st);
    if (r) return r;
    if (fobj_dev) {
      k_eval_scatter<<<(count + 127) / 128, 128, 0, st>>>(W.w.c, W.w.a, count, fobj_dev, nullptr, b->D + 2);
      ++h->launches;
    }
    k_sg_npb<<<(S + 127) / 128, 128, 0, st>>>(g, poff_dev);
    k_sg_unpack_x<<<dim3(mr, mr, S), 256, 0, st>>>(W.cb, g, g.BF, 1);                      // X_B incl. its augmented row
    k_sg_build_uf<<<dim3(nc, mb, S), 256, 0, st>>>(g, g.KUF, pcoords_dev, poff_dev);        // K(Z, X*)
    h->launches += 3;
    sg_gemm<false, true>(h, W, g.XLF, g.KUF, g.AP, g.mb, g.npb, g.mb, mb, nc, 1, 0, st);    // t1 = X_L Kus
    sg_gemm<false, true>(h, W, g.BF, g.AP, g.KUF, g.mb1, g.npb, g.mb, mr, nc, 1, 0, st);    // t2 (+ row M: -mean)
    k_sg_pred_out<<<dim3(nc, S), 256, 0, st>>>(W.cb, g, g.AP, g.KUF, poff_dev, fmean_dev, fvar_dev, yvar_dev);
    ++h->launches;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  return check_sync_timeouts(h);
}
