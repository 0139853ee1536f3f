"""Unconstrained L-BFGS-B restated as an ask/tell state machine.  TEST INFRASTRUCTURE ONLY.

Follows scipy.optimize's L-BFGS-B (Byrd-Lu-Nocedal-Zhu v3.0 `mainlb`, `lnsrlb`,
`dcsrch`, `dcstep`) for the case the reference uses (row P1 of SURVEY.md section 8a,
GPSat/models/gpflow_models.py:317-321: no bounds, scipy defaults m=10,
ftol=2.22e-9 (factr=1e7), gtol(pgtol)=1e-5, maxls=20, maxfun=15000, maxiter from
the caller).  With no bounds the generalized Cauchy point / subspace
minimisation collapse to the L-BFGS two-loop direction with H0 = I/theta,
theta = y'y / s'y; everything else (first-step length 1/|d|, Moré-Thuente line
search with ftol=1e-3, gtol=0.9, xtol=0.1, update-skip rule s'y <= eps*(-g'd),
line-search failure -> memory reset -> abnormal termination) is kept.

The CUDA batched optimiser (gpsat_b200/csrc/lbfgs.cuh) is a device port of this
file; tests compare the two and pin this file against scipy itself.
"""
from __future__ import annotations

import math
import numpy as np

EPSMCH = 2.220446049250313e-16
FTOL_LS, GTOL_LS, XTOL_LS = 1e-3, 0.9, 0.1
XTRAPL, XTRAPU = 1.1, 4.0
BIG = 1e10

# status codes (shared with the device code)
RUNNING = 0
CONV_PGTOL = 1        # CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL   (success)
CONV_FTOL = 2         # CONVERGENCE: REL_REDUCTION_OF_F <= FACTR*EPSMCH    (success)
STOP_MAXITER = 3      # STOP: TOTAL NO. of ITERATIONS REACHED LIMIT        (fail)
STOP_MAXFUN = 4       # STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT (fail)
ABNORMAL = 5          # ABNORMAL_TERMINATION_IN_LNSRCH                     (fail)


def dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx):
            stpf = stpc
        else:
            stpf = stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = min(stpmax, stpf)
            stpf = max(stpmin, stpf)
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * math.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


class Lbfgs:
    """ask/tell optimiser: evaluate (f, g) at ``self.x`` then call ``tell(f, g)``."""

    def __init__(self, x0, m=10, factr=1e7, pgtol=1e-5, maxiter=10_000, maxfun=15_000, maxls=20):
        self.n = len(x0)
        self.m, self.factr, self.pgtol = m, factr, pgtol
        self.maxiter, self.maxfun, self.maxls = maxiter, maxfun, maxls
        self.x = np.array(x0, dtype=np.float64)
        self.status = RUNNING
        self.phase = 0            # 0: waiting for f(x0); 1: inside line search
        self.nit = 0
        self.nfev = 0
        self.S = np.zeros((m, self.n))
        self.Y = np.zeros((m, self.n))
        self.col = 0
        self.head = 0             # index of the oldest pair
        self.theta = 1.0
        self.f = None
        self.g = None

    # ---- two-loop direction with H0 = I/theta over the stored pairs ----
    def _direction(self):
        q = -self.g.copy()
        if self.col == 0:
            return q
        idx = [(self.head + i) % self.m for i in range(self.col)]  # oldest..newest
        al = np.zeros(self.col)
        for k in reversed(range(self.col)):
            s, y = self.S[idx[k]], self.Y[idx[k]]
            al[k] = (s @ q) / (y @ s)
            q = q - al[k] * y
        q = q / self.theta
        for k in range(self.col):
            s, y = self.S[idx[k]], self.Y[idx[k]]
            be = (y @ q) / (y @ s)
            q = q + (al[k] - be) * s
        return q

    def _start_iteration(self):
        """label 222 of mainlb: direction + line-search initialisation; sets next trial x."""
        while True:
            self.d = self._direction()
            self.t = self.x.copy()          # x at start of the line search
            self.fold = self.f
            self.r = self.g.copy()          # old gradient
            dnorm = math.sqrt(float(self.d @ self.d))
            self.stpmx = BIG
            self.stp = min(1.0 / dnorm, self.stpmx) if (self.nit == 0) else 1.0
            self.ifun = 0
            self.iback = 0
            gd = float(self.g @ self.d)
            self.gdold = gd
            if gd >= 0.0:
                # ascent direction: info = -4 -> restart without memory, or abnormal
                if self.col == 0:
                    self.status = ABNORMAL
                    return
                self._reset_memory()
                continue
            # dcsrch START
            self.brackt = False
            self.stage = 1
            self.finit = self.f
            self.ginit = gd
            self.gtest = FTOL_LS * gd
            self.width = self.stpmx - 0.0
            self.width1 = self.width / 0.5
            self.stx, self.fx, self.gx = 0.0, self.finit, self.ginit
            self.sty, self.fy, self.gy = 0.0, self.finit, self.ginit
            self.stmin = 0.0
            self.stmax = self.stp + XTRAPU * self.stp
            self._issue_trial()
            return

    def _issue_trial(self):
        self.ifun += 1
        self.iback = self.ifun - 1
        if self.iback >= self.maxls:
            self._linesearch_failed()
            return
        if self.stp == 1.0:
            self.x = self.t + self.d
        else:
            self.x = self.stp * self.d + self.t
        self.phase = 1

    def _reset_memory(self):
        self.col, self.head, self.theta = 0, 0, 1.0

    def _linesearch_failed(self):
        # restore previous iterate (mainlb after lnsrlb with info != 0 or iback >= maxls)
        self.x = self.t.copy()
        self.f = self.fold
        self.g = self.r.copy()
        if self.col == 0:
            self.status = ABNORMAL
            return
        self._reset_memory()
        self._start_iteration()

    def tell(self, f, g):
        assert self.status == RUNNING
        self.nfev += 1
        g = np.asarray(g, dtype=np.float64)
        if self.phase == 0:
            self.f, self.g = float(f), g.copy()
            if np.max(np.abs(self.g)) <= self.pgtol:
                self.status = CONV_PGTOL
                return
            self._start_iteration()
            return
        # ---- inside the line search: dcsrch with a new (f, g) at stp ----
        f = float(f)
        if not math.isfinite(f):
            # deviation (the reference aborts on a failed Cholesky): treat as a very
            # bad point -- bisect towards the best step so far.
            self.stp = 0.5 * (self.stx + self.stp)
            self._issue_trial()
            return
        gd = float(g @ self.d)
        stp = self.stp
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and gd >= 0.0:
            self.stage = 2
        done = False
        if self.brackt and (stp <= self.stmin or stp >= self.stmax):
            done = True
        if self.brackt and self.stmax - self.stmin <= XTOL_LS * self.stmax:
            done = True
        if stp == self.stpmx and f <= ftest and gd <= self.gtest:
            done = True
        if stp == 0.0 and (f > ftest or gd >= self.gtest):
            done = True
        if f <= ftest and abs(gd) <= GTOL_LS * (-self.ginit):
            done = True
        if done:
            self._accept(f, g, gd)
            return
        if self.stage == 1 and f <= self.fx and f > ftest:
            fm = f - stp * self.gtest
            fxm = self.fx - self.stx * self.gtest
            fym = self.fy - self.sty * self.gtest
            gm = gd - self.gtest
            gxm = self.gx - self.gtest
            gym = self.gy - self.gtest
            (self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt) = dcstep(
                self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm, self.brackt,
                self.stmin, self.stmax)
            self.fx = fxm + self.stx * self.gtest
            self.fy = fym + self.sty * self.gtest
            self.gx = gxm + self.gtest
            self.gy = gym + self.gtest
        else:
            (self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, self.brackt) = dcstep(
                self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, gd, self.brackt,
                self.stmin, self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= 0.66 * self.width1:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.width1 = self.width
            self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin = min(self.stx, self.sty)
            self.stmax = max(self.stx, self.sty)
        else:
            self.stmin = stp + XTRAPL * (stp - self.stx)
            self.stmax = stp + XTRAPU * (stp - self.stx)
        stp = max(stp, 0.0)
        stp = min(stp, self.stpmx)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or \
                (self.brackt and self.stmax - self.stmin <= XTOL_LS * self.stmax):
            stp = self.stx
        self.stp = stp
        self._issue_trial()

    def _accept(self, f, g, gd):
        """NEW_X: label 777 of mainlb."""
        self.f, self.g = f, g.copy()
        self.nit += 1
        sbgnrm = float(np.max(np.abs(self.g)))
        # scipy's python loop checks maxiter / maxfun at NEW_X before re-entering setulb
        if self.nit >= self.maxiter:
            self.status = STOP_MAXITER
            return
        if self.nfev > self.maxfun:
            self.status = STOP_MAXFUN
            return
        if sbgnrm <= self.pgtol:
            self.status = CONV_PGTOL
            return
        ddum = max(abs(self.fold), abs(self.f), 1.0)
        if (self.fold - self.f) <= EPSMCH * self.factr * ddum:
            self.status = CONV_FTOL
            return
        # BFGS pair
        y = self.g - self.r
        rr = float(y @ y)
        if self.stp == 1.0:
            dr = gd - self.gdold
            ddum = -self.gdold
            s = self.d
        else:
            dr = (gd - self.gdold) * self.stp
            s = self.stp * self.d
            ddum = -self.gdold * self.stp
        if dr > EPSMCH * ddum:
            if self.col < self.m:
                slot = (self.head + self.col) % self.m
                self.col += 1
            else:
                slot = self.head
                self.head = (self.head + 1) % self.m
            self.S[slot] = s
            self.Y[slot] = y
            self.theta = rr / dr
        self._start_iteration()


def minimize_lbfgs(fun, x0, maxiter=10_000, **kw):
    opt = Lbfgs(x0, maxiter=maxiter, **kw)
    while opt.status == RUNNING:
        f, g = fun(opt.x)
        opt.tell(f, g)
    return {"x": opt.x, "fun": opt.f, "jac": opt.g, "nit": opt.nit, "nfev": opt.nfev,
            "status": opt.status, "success": opt.status in (CONV_PGTOL, CONV_FTOL)}
