// gpsat_b200: device-resident batched L-BFGS (one optimiser state per expert slot).
//
// Device port of oracle/lbfgs.py, which restates scipy's unconstrained L-BFGS-B as reached
// from GPSat/models/gpflow_models.py:317-321 (gpflow.optimizers.Scipy -> scipy L-BFGS-B with
// maxcor=10, ftol=2.22e-9, gtol=1e-5, maxls=20, maxfun=15000, maxiter from the caller).
// ask/tell form: the engine evaluates (f, g) at st.x for every running slot (one batched
// "round" of factorisation kernels), then one thread per slot calls lbfgs_tell().
#pragma once
#include "common.cuh"
#include <math.h>
#define GPSAT_HD __host__ __device__

namespace gpsat {

constexpr int LB_M = 10;

enum LbStatus { LB_RUNNING = 0, LB_CONV_PGTOL = 1, LB_CONV_FTOL = 2, LB_STOP_MAXITER = 3,
                LB_STOP_MAXFUN = 4, LB_ABNORMAL = 5 };

struct LbfgsOpts {
  int m;          // history pairs (<= LB_M)
  int maxiter;
  int maxfun;
  int maxls;
  double factr;   // ftol / eps
  double pgtol;
};

struct LbfgsState {
  int n, status, phase, nit, nfev, col, head, ifun, iback, brackt, stage, pad_;
  double theta, f, fold, stp, stpmx, gdold, finit, ginit, gtest, width, width1;
  double stx, fx, gx, sty, fy, gy, stmin, stmax;
  double x[MAXP], g[MAXP], d[MAXP], t[MAXP], r[MAXP];
  double S[LB_M][MAXP], Y[LB_M][MAXP];
};

#define LB_EPSMCH 2.220446049250313e-16

GPSAT_HD inline void lb_init(LbfgsState& st, const double* x0, int n) {
  st.n = n;
  st.status = LB_RUNNING;
  st.phase = 0;
  st.nit = 0;
  st.nfev = 0;
  st.col = 0;
  st.head = 0;
  st.theta = 1.0;
  st.f = 0.0;
  for (int i = 0; i < n; ++i) st.x[i] = x0[i];
}

GPSAT_HD inline double lb_dot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

GPSAT_HD inline void lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy,
                                 double& dy, double& stp, double fp, double dp, int& brackt,
                                 double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    double p = (gamma - dx) + theta;
    double q = ((gamma - dx) + gamma) + dp;
    double r = p / q;
    double stpc = stx + r * (stp - stx);
    double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    double p = (gamma - dp) + theta;
    double q = ((gamma - dp) + gamma) + dx;
    double r = p / q;
    double stpc = stp + r * (stx - stp);
    double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    double p = (gamma - dp) + theta;
    double q = (gamma + (dx - dp)) + gamma;
    double r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      double p = (gamma - dp) + theta;
      double q = ((gamma - dp) + gamma) + dy;
      double r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// two-loop recursion, H0 = I/theta
GPSAT_HD inline void lb_direction(LbfgsState& st, int m) {
  const int n = st.n;
  double q[MAXP];
  for (int i = 0; i < n; ++i) q[i] = -st.g[i];
  if (st.col > 0) {
    double al[LB_M];
    for (int k = st.col - 1; k >= 0; --k) {
      const int id = (st.head + k) % m;
      const double ys = lb_dot(st.Y[id], st.S[id], n);
      al[k] = lb_dot(st.S[id], q, n) / ys;
      for (int i = 0; i < n; ++i) q[i] -= al[k] * st.Y[id][i];
    }
    for (int i = 0; i < n; ++i) q[i] /= st.theta;
    for (int k = 0; k < st.col; ++k) {
      const int id = (st.head + k) % m;
      const double ys = lb_dot(st.Y[id], st.S[id], n);
      const double be = lb_dot(st.Y[id], q, n) / ys;
      for (int i = 0; i < n; ++i) q[i] += (al[k] - be) * st.S[id][i];
    }
  }
  for (int i = 0; i < n; ++i) st.d[i] = q[i];
}

GPSAT_HD inline void lb_linesearch_failed(LbfgsState& st, const LbfgsOpts& o);

GPSAT_HD inline void lb_issue_trial(LbfgsState& st, const LbfgsOpts& o) {
  st.ifun += 1;
  st.iback = st.ifun - 1;
  if (st.iback >= o.maxls) {
    lb_linesearch_failed(st, o);
    return;
  }
  if (st.stp == 1.0) {
    for (int i = 0; i < st.n; ++i) st.x[i] = st.t[i] + st.d[i];
  } else {
    for (int i = 0; i < st.n; ++i) st.x[i] = st.stp * st.d[i] + st.t[i];
  }
  st.phase = 1;
}

// label 222 of mainlb: direction, line-search start, first trial point.
// `depth` guards the (bounded) restart recursion.
GPSAT_HD inline void lb_start_iteration(LbfgsState& st, const LbfgsOpts& o) {
  for (int attempt = 0; attempt < 3; ++attempt) {
    lb_direction(st, o.m);
    for (int i = 0; i < st.n; ++i) { st.t[i] = st.x[i]; st.r[i] = st.g[i]; }
    st.fold = st.f;
    const double dnorm = sqrt(lb_dot(st.d, st.d, st.n));
    st.stpmx = 1e10;
    st.stp = (st.nit == 0) ? fmin(1.0 / dnorm, st.stpmx) : 1.0;
    st.ifun = 0;
    st.iback = 0;
    const double gd = lb_dot(st.g, st.d, st.n);
    st.gdold = gd;
    if (!(gd < 0.0)) {  // ascent (or NaN) direction
      if (st.col == 0) { st.status = LB_ABNORMAL; return; }
      st.col = 0; st.head = 0; st.theta = 1.0;
      continue;
    }
    st.brackt = 0;
    st.stage = 1;
    st.finit = st.f;
    st.ginit = gd;
    st.gtest = 1e-3 * gd;
    st.width = st.stpmx;
    st.width1 = st.width / 0.5;
    st.stx = 0.0; st.fx = st.finit; st.gx = st.ginit;
    st.sty = 0.0; st.fy = st.finit; st.gy = st.ginit;
    st.stmin = 0.0;
    st.stmax = st.stp + 4.0 * st.stp;
    lb_issue_trial(st, o);
    return;
  }
  st.status = LB_ABNORMAL;
}

GPSAT_HD inline void lb_linesearch_failed(LbfgsState& st, const LbfgsOpts& o) {
  for (int i = 0; i < st.n; ++i) { st.x[i] = st.t[i]; st.g[i] = st.r[i]; }
  st.f = st.fold;
  if (st.col == 0) { st.status = LB_ABNORMAL; return; }
  st.col = 0; st.head = 0; st.theta = 1.0;
  // restart from the restored iterate with an empty memory (steepest descent, unit step rule
  // of a non-first iteration).  One level only: a second failure has col == 0 -> ABNORMAL.
  lb_direction(st, o.m);
  st.fold = st.f;
  const double gd = lb_dot(st.g, st.d, st.n);
  st.gdold = gd;
  if (!(gd < 0.0)) { st.status = LB_ABNORMAL; return; }
  st.stpmx = 1e10;
  st.stp = (st.nit == 0) ? fmin(1.0 / sqrt(lb_dot(st.d, st.d, st.n)), st.stpmx) : 1.0;
  st.ifun = 0; st.iback = 0;
  st.brackt = 0; st.stage = 1;
  st.finit = st.f; st.ginit = gd; st.gtest = 1e-3 * gd;
  st.width = st.stpmx; st.width1 = st.width / 0.5;
  st.stx = 0.0; st.fx = st.finit; st.gx = st.ginit;
  st.sty = 0.0; st.fy = st.finit; st.gy = st.ginit;
  st.stmin = 0.0; st.stmax = st.stp + 4.0 * st.stp;
  st.ifun = 1; st.iback = 0;
  if (st.stp == 1.0) { for (int i = 0; i < st.n; ++i) st.x[i] = st.t[i] + st.d[i]; }
  else { for (int i = 0; i < st.n; ++i) st.x[i] = st.stp * st.d[i] + st.t[i]; }
  st.phase = 1;
}

GPSAT_HD inline void lb_accept(LbfgsState& st, const LbfgsOpts& o, double f, const double* g, double gd) {
  st.f = f;
  for (int i = 0; i < st.n; ++i) st.g[i] = g[i];
  st.nit += 1;
  double sbgnrm = 0.0;
  for (int i = 0; i < st.n; ++i) sbgnrm = fmax(sbgnrm, fabs(g[i]));
  if (st.nit >= o.maxiter) { st.status = LB_STOP_MAXITER; return; }
  if (st.nfev > o.maxfun) { st.status = LB_STOP_MAXFUN; return; }
  if (sbgnrm <= o.pgtol) { st.status = LB_CONV_PGTOL; return; }
  double ddum = fmax(fabs(st.fold), fmax(fabs(st.f), 1.0));
  if ((st.fold - st.f) <= LB_EPSMCH * o.factr * ddum) { st.status = LB_CONV_FTOL; return; }
  double y[MAXP];
  for (int i = 0; i < st.n; ++i) y[i] = st.g[i] - st.r[i];
  const double rr = lb_dot(y, y, st.n);
  double dr, sc;
  if (st.stp == 1.0) { dr = gd - st.gdold; ddum = -st.gdold; sc = 1.0; }
  else { dr = (gd - st.gdold) * st.stp; ddum = -st.gdold * st.stp; sc = st.stp; }
  if (dr > LB_EPSMCH * ddum) {
    int slot;
    if (st.col < o.m) { slot = (st.head + st.col) % o.m; st.col += 1; }
    else { slot = st.head; st.head = (st.head + 1) % o.m; }
    for (int i = 0; i < st.n; ++i) { st.S[slot][i] = sc * st.d[i]; st.Y[slot][i] = y[i]; }
    st.theta = rr / dr;
  }
  lb_start_iteration(st, o);
}

// feed (f, g) evaluated at st.x; on return either st.status != RUNNING or st.x is the next point
GPSAT_HD inline void lbfgs_tell(LbfgsState& st, const LbfgsOpts& o, double f, const double* g) {
  st.nfev += 1;
  if (st.phase == 0) {
    st.f = f;
    double sb = 0.0;
    for (int i = 0; i < st.n; ++i) { st.g[i] = g[i]; sb = fmax(sb, fabs(g[i])); }
    if (!isfinite(f)) { st.status = LB_ABNORMAL; return; }
    if (sb <= o.pgtol) { st.status = LB_CONV_PGTOL; return; }
    lb_start_iteration(st, o);
    return;
  }
  if (!isfinite(f)) {
    // deviation from the reference (which aborts on a failed Cholesky): bisect towards stx
    st.stp = 0.5 * (st.stx + st.stp);
    lb_issue_trial(st, o);
    return;
  }
  const double gd = lb_dot(g, st.d, st.n);
  double stp = st.stp;
  const double ftest = st.finit + stp * st.gtest;
  if (st.stage == 1 && f <= ftest && gd >= 0.0) st.stage = 2;
  bool done = false;
  if (st.brackt && (stp <= st.stmin || stp >= st.stmax)) done = true;
  if (st.brackt && st.stmax - st.stmin <= 0.1 * st.stmax) done = true;
  if (stp == st.stpmx && f <= ftest && gd <= st.gtest) done = true;
  if (stp == 0.0 && (f > ftest || gd >= st.gtest)) done = true;
  if (f <= ftest && fabs(gd) <= 0.9 * (-st.ginit)) done = true;
  if (done) { lb_accept(st, o, f, g, gd); return; }
  if (st.stage == 1 && f <= st.fx && f > ftest) {
    double fm = f - stp * st.gtest, fxm = st.fx - st.stx * st.gtest, fym = st.fy - st.sty * st.gtest;
    double gm = gd - st.gtest, gxm = st.gx - st.gtest, gym = st.gy - st.gtest;
    lb_dcstep(st.stx, fxm, gxm, st.sty, fym, gym, stp, fm, gm, st.brackt, st.stmin, st.stmax);
    st.fx = fxm + st.stx * st.gtest;
    st.fy = fym + st.sty * st.gtest;
    st.gx = gxm + st.gtest;
    st.gy = gym + st.gtest;
  } else {
    lb_dcstep(st.stx, st.fx, st.gx, st.sty, st.fy, st.gy, stp, f, gd, st.brackt, st.stmin, st.stmax);
  }
  if (st.brackt) {
    if (fabs(st.sty - st.stx) >= 0.66 * st.width1) stp = st.stx + 0.5 * (st.sty - st.stx);
    st.width1 = st.width;
    st.width = fabs(st.sty - st.stx);
  }
  if (st.brackt) {
    st.stmin = fmin(st.stx, st.sty);
    st.stmax = fmax(st.stx, st.sty);
  } else {
    st.stmin = stp + 1.1 * (stp - st.stx);
    st.stmax = stp + 4.0 * (stp - st.stx);
  }
  stp = fmax(stp, 0.0);
  stp = fmin(stp, st.stpmx);
  if ((st.brackt && (stp <= st.stmin || stp >= st.stmax)) ||
      (st.brackt && st.stmax - st.stmin <= 0.1 * st.stmax))
    stp = st.stx;
  st.stp = stp;
  lb_issue_trial(st, o);
}

}  // namespace gpsat
