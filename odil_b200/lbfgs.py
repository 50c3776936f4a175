"""
Device-resident L-BFGS for the `lbfgsb` optimizer name (unconstrained problems: ODIL never sets bounds).

The reference calls SciPy's `fmin_l_bfgs_b` on a flat fp64 HOST vector and moves the whole state across
PCIe twice per evaluation (src/odil/optimizer.py:54-117).  Here the unknown vector, the gradient and the
2m history vectors stay in HBM (fp64, like the reference's host vector); per iteration the device does
ONE pass over the history for all inner products [S Y]^T g (odil_b200_multi_dot) and ONE pass for the
direction d = -H g (odil_b200_multi_axpy) -- the compact (Byrd-Nocedal-Schnabel) form that L-BFGS-B uses
internally -- and only O(m^2) scalars visit the host.

To make the same step-length decisions as the reference, the control flow follows L-BFGS-B 3.0 as shipped
in SciPy (`lbfgsb.f`: mainlb / lnsrlb / matupd) restricted to the unconstrained case, and the line search
is a transcription of MINPACK-2 `dcsrch` / `dcstep` (More-Thuente) with L-BFGS-B's constants
ftol=1e-3, gtol=0.9, xtol=0.1, first step min(1/|d|, stpmx), later steps 1.
"""
import numpy as np
import torch

from . import native

EPSMCH = np.finfo(np.float64).eps
BIG = 1e10
FTOL, GTOL, XTOL = 1e-3, 0.9, 0.1


class _Search:
    """State of one dcsrch line search (MINPACK-2)."""
    XTRAPL, XTRAPU = 1.1, 4.0

    def start(self, stp, f, g, stpmin, stpmax):
        if stp < stpmin or stp > stpmax or g >= 0 or stpmax < stpmin:
            return "ERROR"
        self.brackt = False
        self.stage = 1
        self.finit, self.ginit = f, g
        self.gtest = FTOL * g
        self.width = stpmax - stpmin
        self.width1 = 2 * self.width
        self.stx, self.fx, self.gx = 0.0, f, g
        self.sty, self.fy, self.gy = 0.0, f, g
        self.stmin = 0.0
        self.stmax = stp + self.XTRAPU * stp
        self.stpmin, self.stpmax = stpmin, stpmax
        return "FG"

    def step(self, stp, f, g):
        """One dcsrch call after an evaluation at `stp`. Returns (task, new_stp)."""
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and g >= 0:
            self.stage = 2
        task = None
        if self.brackt and (stp <= self.stmin or stp >= self.stmax):
            task = "WARNING: ROUNDING ERRORS PREVENT PROGRESS"
        if self.brackt and self.stmax - self.stmin <= XTOL * self.stmax:
            task = "WARNING: XTOL TEST SATISFIED"
        if stp == self.stpmax and f <= ftest and g <= self.gtest:
            task = "WARNING: STP = STPMAX"
        if stp == self.stpmin and (f > ftest or g >= self.gtest):
            task = "WARNING: STP = STPMIN"
        if f <= ftest and abs(g) <= GTOL * (-self.ginit):
            task = "CONVERGENCE"
        if task is not None:
            return task, stp
        if self.stage == 1 and f <= self.fx and f > ftest:
            gt = self.gtest
            fm, fxm, fym = f - stp * gt, self.fx - self.stx * gt, self.fy - self.sty * gt
            gm, gxm, gym = g - gt, self.gx - gt, self.gy - gt
            (self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt) = _dcstep(
                self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm, self.brackt, self.stmin, self.stmax)
            self.fx, self.fy = fxm + self.stx * gt, fym + self.sty * gt
            self.gx, self.gy = gxm + gt, gym + gt
        else:
            (self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, self.brackt) = _dcstep(
                self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, g, self.brackt, self.stmin, self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= 0.66 * self.width1:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.width1 = self.width
            self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin, self.stmax = min(self.stx, self.sty), max(self.stx, self.sty)
        else:
            self.stmin = stp + self.XTRAPL * (stp - self.stx)
            self.stmax = stp + self.XTRAPU * (stp - self.stx)
        stp = min(max(stp, self.stpmin), self.stpmax)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or \
                (self.brackt and self.stmax - self.stmin <= XTOL * self.stmax):
            stp = self.stx
        return "FG", stp


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    """MINPACK-2 dcstep: safeguarded cubic/quadratic step and interval update."""
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
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
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
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
            gamma = s * np.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
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


class _History:
    """S and Y as rows of one device matrix V (pair p occupies rows 2p and 2p+1; a ring: the oldest pair's rows are
    reused) plus the small host matrices S^T Y and Y^T Y kept in age order.  Only the rows in use are streamed: while
    the history is filling, the two passes per iteration cost 2 k vectors each, not 2 m."""

    def __init__(self, m, n, device):
        self.m, self.n = m, n
        self.V = torch.zeros((2 * m, n), dtype=torch.float64, device=device)  # zeros: multi_dot reads all rows
        self.rows = []               # physical row of each stored pair, oldest first
        self.sy = np.zeros((m, m))   # sy[i, j] = s_i . y_j   (age order)
        self.yy = np.zeros((m, m))
        self.theta = 1.0
        self._out = torch.zeros(2 * m, dtype=torch.float64, device=device)
        self._coef = torch.zeros(2 * m, dtype=torch.float64, device=device)

    @property
    def col(self):
        return len(self.rows)

    def reset(self):
        self.rows = []
        self.theta = 1.0

    def dots(self, vec):
        """(S^T vec, Y^T vec) for the stored pairs in age order: one device pass over V, one sync."""
        native.multi_dot(self.V, 2 * len(self.rows), vec, self._out)   # physical rows 0 .. 2 * col - 1 are in use
        h = self._out.cpu().numpy()
        idx = np.asarray(self.rows, dtype=int)
        return h[2 * idx].copy(), h[2 * idx + 1].copy()

    def push(self, s, y, sty, yty, Sty, Yty):
        """Appends the pair (s, y); Sty = S_old^T y, Yty = Y_old^T y (host, age order), sty = s.y, yty = y.y."""
        m = self.m
        if len(self.rows) == m:  # drop the oldest pair, reuse its rows
            row = self.rows.pop(0)
            self.sy[:m - 1, :m - 1] = self.sy[1:, 1:]
            self.yy[:m - 1, :m - 1] = self.yy[1:, 1:]
            Sty, Yty = Sty[1:], Yty[1:]
        else:
            row = len(self.rows)
        k = len(self.rows)
        self.V[2 * row].copy_(s)
        self.V[2 * row + 1].copy_(y)
        self.rows.append(row)
        self.sy[:k, k] = Sty
        self.sy[k, :k] = 0.0
        self.sy[k, k] = sty
        self.yy[:k, k] = Yty
        self.yy[k, :k] = Yty
        self.yy[k, k] = yty
        self.theta = yty / sty

    def direction(self, g, d, Stg, Ytg):
        """d = -H g with the compact inverse form; Stg, Ytg = S^T g, Y^T g (host arrays, age order)."""
        k = len(self.rows)
        gam = 1.0 / self.theta
        R = np.triu(self.sy[:k, :k])
        D = np.diag(np.diag(self.sy[:k, :k]))
        p1, p2 = Stg, gam * Ytg
        Rinv_p1 = np.linalg.solve(R, p1)
        u1 = np.linalg.solve(R.T, (D + gam * self.yy[:k, :k]) @ Rinv_p1 - p2)
        u2 = -Rinv_p1
        coef = np.zeros(2 * self.m)
        idx = np.asarray(self.rows, dtype=int)
        coef[2 * idx] = -u1
        coef[2 * idx + 1] = -gam * u2
        self._coef.copy_(torch.from_numpy(coef))
        native.multi_axpy(self.V, 2 * k, self._coef, -gam, g, d)


def _dot(a, b, out):
    native.dot(a, b, out)
    return out.item()


def _maxabs(g):
    """max |g_i| in one pass over g (g.abs().max() makes a temporary and reads it back)."""
    return torch.linalg.vector_norm(g, ord=float("inf")).item()


def minimize(func, x0, m=50, maxiter=15000, maxls=20, pgtol=1e-5, factr=1e7, callback=None):
    """
    func(x) -> (f, g): x, g flat fp64 device tensors (g may be overwritten by the caller on the next call).
    Returns (x, f, info) with SciPy's info keys: warnflag, task, funcalls, nit.
    """
    n = x0.numel()
    dev = x0.device
    x = x0.clone()
    hist = _History(m, n, dev)
    t = torch.empty_like(x)   # previous iterate
    r = torch.empty_like(x)   # previous gradient, then y
    d = torch.empty_like(x)
    sc = torch.zeros(1, dtype=torch.float64, device=dev)
    f, g = func(x)
    g = g.clone()
    nfev, nit = 1, 0
    task, warnflag = None, 0
    sbgnrm = _maxabs(g)
    if sbgnrm <= pgtol:
        return x, f, dict(warnflag=0, task="CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL", funcalls=nfev, nit=0)
    Stg = Ytg = None
    while True:
        # ---- search direction ------------------------------------------------------------------
        if hist.col == 0:
            native.axpby(-1.0, g, 0.0, d)
        else:
            if Stg is None:
                Stg, Ytg = hist.dots(g)
            hist.direction(g, d, Stg, Ytg)
        # ---- line search (lnsrlb) ----------------------------------------------------------------
        dnorm = np.sqrt(_dot(d, d, sc))
        stp = min(1.0 / dnorm, BIG) if nit == 0 else 1.0
        t.copy_(x)
        g, r = r, g   # r = previous gradient (kept for y = g_new - g_old and for a failed search); g is overwritten below
        fold = f
        gd = _dot(r, d, sc)
        gdold = gd
        ls = _Search()
        info = 0
        iback = 0
        if gd >= 0 or ls.start(stp, f, gd, 0.0, BIG) == "ERROR":
            info = -4
        lstask = "FG"
        while info == 0:
            # evaluate at x = t + stp * d
            x.copy_(t)
            native.axpby(stp, d, 1.0, x)
            f, gnew = func(x)
            g.copy_(gnew)
            nfev += 1
            iback += 1
            gd = _dot(g, d, sc)
            lstask, stp_next = ls.step(stp, f, gd)
            if lstask != "FG":
                break
            if iback >= maxls:
                break
            stp = stp_next
        if info != 0 or (lstask == "FG" and iback >= maxls):
            # restore the previous iterate; restart from steepest descent if there is history to drop
            x.copy_(t)
            g.copy_(r)
            f = fold
            if hist.col == 0:
                task, warnflag = "ABNORMAL_TERMINATION_IN_LNSRCH", 2
                nit += 1
                break
            hist.reset()
            Stg = Ytg = None
            continue
        # ---- new iterate -------------------------------------------------------------------------
        nit += 1
        if callback is not None:
            callback(x)
        sbgnrm = _maxabs(g)
        if sbgnrm <= pgtol:
            task = "CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL"
            break
        if (fold - f) <= EPSMCH * factr * max(abs(fold), abs(f), 1.0):
            task = "CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH"
            break
        if nit >= maxiter:
            task, warnflag = "STOP: TOTAL NO. of ITERATIONS REACHED LIMIT", 1
            break
        # ---- update the limited-memory matrices (matupd) -----------------------------------------
        Stg_old, Ytg_old = Stg, Ytg
        Stg_new, Ytg_new = (hist.dots(g) if hist.col > 0 else (np.zeros(0), np.zeros(0)))
        native.axpby(1.0, g, -1.0, r)            # r = g - g_old = y
        if stp == 1.0:
            dr, ddum = gd - gdold, -gdold
        else:
            dr, ddum = (gd - gdold) * stp, -gdold * stp
            native.axpby(0.0, d, stp, d)          # d *= stp  (= s)
        rr = _dot(r, r, sc)
        if dr <= EPSMCH * ddum:
            Stg, Ytg = Stg_new, Ytg_new           # skip the update, keep the history
            continue
        if hist.col > 0 and Stg_old is not None:
            Sty, Yty = Stg_new - Stg_old, Ytg_new - Ytg_old   # S^T y = S^T g_new - S^T g_old
        elif hist.col > 0:
            Sty, Yty = hist.dots(r)
        else:
            Sty, Yty = np.zeros(0), np.zeros(0)
        was_full = hist.col == hist.m
        hist.push(d, r, dr, rr, Sty, Yty)
        # S^T g and Y^T g of the UPDATED history without another pass over it: the old pairs' products with the new
        # gradient were just computed (Stg_new, Ytg_new; the oldest entry leaves with its pair), the new pair adds
        # s.g = stp * (d.g) -- known from the line search -- and y.g (one dot product)
        sg = gd * stp if stp != 1.0 else gd
        yg = _dot(r, g, sc)
        drop = 1 if was_full else 0
        Stg = np.concatenate([Stg_new[drop:], [sg]])
        Ytg = np.concatenate([Ytg_new[drop:], [yg]])
    return x, f, dict(warnflag=warnflag, task=task, funcalls=nfev, nit=nit)
