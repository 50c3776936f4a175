"""
CPU oracle for the ODIL residual-and-gradient hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Plain-NumPy restatement of the reference algorithm (cselab/odil @ a794c01).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product package `odil_b200` never does (it fails loudly when the CUDA library
is missing).

Parity pinning: the reference's own tests hold no stored numbers for this path,
so the oracle is pinned against `tests/golden/*.npz`, which were produced by running the
UNMODIFIED reference `core.py` + example operators here (`tests/golden/make_goldens.py`).
`tests/test_oracle_golden.py` checks every function below against those vectors.

Each function cites the reference lines it restates (paths relative to /root/reference).
All arrays are C-order, axis 0 slowest.  Gradients are written as explicit adjoints (the
reference obtains them by AD: core.py:1100-1101).
"""
import itertools

import numpy as np


# ------------------------------------------------------------------------------------------
# Multigrid transfers
# ------------------------------------------------------------------------------------------
def _padded_take(u, idx_lists, loc):
    """
    Value of the jointly padded array  upad = 2*symmetric(u) - reflect(u)  (core.py:640-643)
    at the padded indices given per axis (values in [-1, n]).  Returns the outer-product gather.
    """
    sym = []
    ref = []
    for ax, q in enumerate(idx_lists):
        n = u.shape[ax]
        q = np.asarray(q)
        sym.append(np.clip(q, 0, n - 1))
        r = q.copy()
        r[q < 0] = 1 if n > 1 else 0
        r[q > n - 1] = n - 2 if n > 1 else 0
        ref.append(r)
    return 2 * u[np.ix_(*sym)] - u[np.ix_(*ref)]


def interp_to_finer(u, loc):
    """
    Restates core.py:606-700 (method "stack"; "conv" is pinned to the same answers).
    Per axis with loc 'c': fine[2i] = (P(i-1) + 3 P(i))/4, fine[2i+1] = (3 P(i) + P(i+1))/4;
    loc 'n': fine[2i] = u[i], fine[2i+1] = (u[i] + u[i+1])/2; loc '.': identity.  The pad P is
    applied jointly over the 'c' axes, so corners are NOT the tensor product of 1-D rules.
    """
    u = np.asarray(u)
    nd = u.ndim
    assert len(loc) == nd
    oshape = tuple({"c": 2 * n, "n": 2 * (n - 1) + 1, ".": n}[l] for n, l in zip(u.shape, loc))
    # Per axis: list of taps (coarse padded index array over fine index, integer weight)
    taps = []
    den = 1
    for n, l, m in zip(u.shape, loc, oshape):
        f = np.arange(m)
        i, a = f // 2, f % 2
        if l == "c":
            taps.append([(np.where(a == 0, i - 1, i + 1), 1), (i, 3)])
            den *= 4
        elif l == "n":
            near = i
            far = np.where(a == 0, i, np.minimum(i + 1, n - 1))
            taps.append([(near, 1), (far, 1)])
            den *= 2
        else:
            taps.append([(f, 1)])
    res = np.zeros(oshape, dtype=u.dtype)
    for combo in itertools.product(*taps):
        w = 1
        shape_w = 1
        for _, wk in combo:
            w *= wk
        res = res + w * _padded_take(u, [c[0] for c in combo], loc)
        del shape_w
    return res / den


def interp_adjoint(g, loc, cshape_field):
    """
    Exact transpose of `interp_to_finer` (what AD of core.py:606-700 yields).  `cshape_field`
    is the coarse ARRAY shape.  Built by scattering through the same taps.
    """
    g = np.asarray(g)
    res = np.zeros(cshape_field, dtype=g.dtype)
    taps = []
    den = 1
    for n, l, m in zip(cshape_field, loc, g.shape):
        f = np.arange(m)
        i, a = f // 2, f % 2
        if l == "c":
            taps.append([(np.where(a == 0, i - 1, i + 1), 1), (i, 3)])
            den *= 4
        elif l == "n":
            taps.append([(i, 1), (np.where(a == 0, i, np.minimum(i + 1, n - 1)), 1)])
            den *= 2
        else:
            taps.append([(f, 1)])
    for combo in itertools.product(*taps):
        w = 1
        for _, wk in combo:
            w *= wk
        sym, ref = [], []
        for ax, (q, _) in enumerate(combo):
            n = cshape_field[ax]
            sym.append(np.clip(q, 0, n - 1))
            r = q.copy()
            r[q < 0] = 1 if n > 1 else 0
            r[q > n - 1] = n - 2 if n > 1 else 0
            ref.append(r)
        np.add.at(res, np.ix_(*sym), 2 * w * g / den)
        np.add.at(res, np.ix_(*ref), -w * g / den)
    return res


def restrict_to_coarser(u, loc):
    """
    Restates core.py:703-755: 'c' = mean of the 2 children per axis ([1,1]/2, stride 2);
    'n' = joint linear-extrapolation pad (2*symmetric - reflect) then [1,2,1]/4, stride 2
    (identity on boundary nodes); '.' = identity.
    """
    u = np.asarray(u)
    taps = []
    for n, l in zip(u.shape, loc):
        if l == "c":
            j = np.arange(n // 2)
            taps.append([(2 * j, 0.5), (2 * j + 1, 0.5)])
        elif l == "n":
            j = np.arange((n - 1) // 2 + 1)
            taps.append([(2 * j - 1, 0.25), (2 * j, 0.5), (2 * j + 1, 0.25)])
        else:
            taps.append([(np.arange(n), 1.0)])
    res = 0
    for combo in itertools.product(*taps):
        w = 1.0
        for _, wk in combo:
            w *= wk
        res = res + w * _padded_take(u, [c[0] for c in combo], loc)
    return res


def mg_synthesize(terms, loc, factors=None):
    """U = t0 f0 + I(t1 f1 + I(...))   (core.py:245-263)."""
    factors = factors or [1] * len(terms)
    res = terms[-1] * factors[-1]
    for t, f in zip(reversed(terms[:-1]), reversed(factors[:-1])):
        res = t * f + interp_to_finer(res, loc)
    return res


def mg_adjoint(gU, shapes, loc, factors=None):
    """g_{t_l} = f_l (I^T)^l gU  -- transpose of `mg_synthesize`."""
    factors = factors or [1] * len(shapes)
    grads = []
    g = gU
    for lvl, shp in enumerate(shapes):
        if lvl > 0:
            g = interp_adjoint(g, loc, shp)
        grads.append(g * factors[lvl])
    return grads


# ------------------------------------------------------------------------------------------
# Region-typed affine stencil  F = A U + c   (the plan the tracer emits)
# ------------------------------------------------------------------------------------------
def region_class_1d(n, r):
    """Class of index i along an axis of size n with region half-width r: 0..r-1 low rows,
    r = interior, r+1..2r high rows."""
    i = np.arange(n)
    cls = np.full(n, r)
    cls[i < r] = i[i < r]
    hi = (n - 1 - i) < r
    cls[hi] = 2 * r - (n - 1 - i[hi])
    return cls


def _coef_fields(shape, table, rr):
    """Expands table[(2r0+1),...,noff] to per-cell coefficient arrays [noff] + shape."""
    cls = [region_class_1d(n, r) for n, r in zip(shape, rr)]
    t = table[np.ix_(*cls)]  # shape + (noff,)
    return np.moveaxis(t, -1, 0)


def stencil_forward(U, offsets, table, rr, const=None):
    """F[x] = sum_o table[class(x), o] * U[(x + off_o) mod N] + c[x]
    (reference: ctx.field = roll(U, -shift), core.py:963, + operator arithmetic)."""
    coef = _coef_fields(U.shape, np.asarray(table), rr)
    F = np.zeros(U.shape, dtype=U.dtype) if const is None else np.array(np.broadcast_to(const, U.shape), dtype=U.dtype)
    axes = tuple(range(U.ndim))
    for o, off in enumerate(offsets):
        F = F + coef[o].astype(U.dtype) * np.roll(U, tuple(-int(s) for s in off), axes)
    return F


def stencil_adjoint(F, offsets, table, rr, scale=1.0):
    """g[x] = scale * sum_o table[class(x - off_o), o] * F[(x - off_o) mod N]  (A^T F)."""
    coef = _coef_fields(F.shape, np.asarray(table), rr)
    g = np.zeros(F.shape, dtype=F.dtype)
    axes = tuple(range(F.ndim))
    for o, off in enumerate(offsets):
        g = g + np.roll(coef[o].astype(F.dtype) * F, tuple(int(s) for s in off), axes)
    return g * F.dtype.type(scale)


# ------------------------------------------------------------------------------------------
# Operators of the example corpus, written out directly (independent of the plan path)
# ------------------------------------------------------------------------------------------
def extrap_quadh(u0, u1, u1p):
    """core.py:1439-1445."""
    return (u0 - 6 * u1 + 8 * u1p) / 3


def poisson_residual(U, rhs, steps):
    """examples/poisson/poisson.py:57-68 (BC), :100-113 (stencil): zero-Dirichlet Poisson."""
    F = -np.asarray(rhs, dtype=U.dtype)
    for ax in range(U.ndim):
        n = U.shape[ax]
        um = np.roll(U, 1, ax)
        up = np.roll(U, -1, ax)
        idx = np.arange(n).reshape([-1 if a == ax else 1 for a in range(U.ndim)])
        zero = U.dtype.type(0)
        qm = np.where(idx == 0, extrap_quadh(up, U, zero), um)
        qp = np.where(idx == n - 1, extrap_quadh(um, U, zero), up)
        F = F + (qp - 2 * U + qm) / steps[ax] ** 2
    return F


def poisson_plan(ndim, steps):
    """The region-typed table equivalent to `poisson_residual`."""
    offsets = [(0,) * ndim]
    for ax in range(ndim):
        for s in (-1, 1):
            offsets.append(tuple(s if a == ax else 0 for a in range(ndim)))
    table = np.zeros((3,) * ndim + (len(offsets),))
    for cls in itertools.product(range(3), repeat=ndim):
        for ax in range(ndim):
            h2 = float(steps[ax]) ** 2
            om, op = 1 + 2 * ax, 2 + 2 * ax
            if cls[ax] == 0:
                table[cls][0] += -4 / h2
                table[cls][op] += (4.0 / 3.0) / h2
            elif cls[ax] == 2:
                table[cls][0] += -4 / h2
                table[cls][om] += (4.0 / 3.0) / h2
            else:
                table[cls][0] += -2 / h2
                table[cls][om] += 1 / h2
                table[cls][op] += 1 / h2
    return offsets, table, (1,) * ndim


def wave_residual(U, dt, dx, left_u, right_u, init_u, init_ut, kimp):
    """examples/wave/wave.py:29-75, (t, x) layout."""
    nt, nx = U.shape
    it = np.arange(nt)[:, None]
    ix = np.arange(nx)[None, :]
    utm = np.roll(U, 1, 0)
    utmm = np.roll(U, 2, 0)
    uxm = np.roll(U, (1, 1), (0, 1))
    uxp = np.roll(U, (1, -1), (0, 1))
    left_utm = np.roll(left_u, 1)[:, None]
    right_utm = np.roll(right_u, 1)[:, None]
    uxm = np.where(ix == 0, extrap_quadh(uxp, utm, left_utm), uxm)
    uxp = np.where(ix == nx - 1, extrap_quadh(uxm, utm, right_utm), uxp)
    u_t_tm = (U - utm) / dt
    u_t_tmm = (utm - utmm) / dt
    u_t_tmm = np.where(it == 1, init_ut[None, :], u_t_tmm)
    u_tt = (u_t_tm - u_t_tmm) / dt
    u_xx = (uxm - 2 * utm + uxp) / dx ** 2
    fu = u_tt - u_xx
    u0 = init_u + 0.5 * dt * init_ut
    return np.where(it == 0, (U - u0[None, :]) * kimp, fu)


def wave2_residual(U, dt, dx, dy, bnd, init_u, init_ut, kimp):
    """The wave operator of examples/wave/wave.py:29-75 carried to two space dimensions, (t, x, y) layout
    (BASELINE configs[2]: u_tt = u_xx + u_yy at level t-1).  `bnd` = dict of Dirichlet data on the four faces,
    each of shape (nt, n_other): xlo, xhi (functions of t, y), ylo, yhi (functions of t, x); init_u, init_ut of
    shape (nx, ny).  Same building blocks as the (t, x) operator: quadratic half-cell extrapolation at the faces
    (core.py:1439-1445), `it == 1` uses init_ut, the `it == 0` row imposes the initial field."""
    nt, nx, ny = U.shape
    it = np.arange(nt)[:, None, None]
    ix = np.arange(nx)[None, :, None]
    iy = np.arange(ny)[None, None, :]
    utm = np.roll(U, 1, 0)
    utmm = np.roll(U, 2, 0)
    uxm, uxp = np.roll(utm, 1, 1), np.roll(utm, -1, 1)
    uym, uyp = np.roll(utm, 1, 2), np.roll(utm, -1, 2)
    prev = lambda a: np.roll(a, 1, 0)  # boundary data at time level t-1
    uxm = np.where(ix == 0, extrap_quadh(uxp, utm, prev(bnd["xlo"])[:, None, :]), uxm)
    uxp = np.where(ix == nx - 1, extrap_quadh(uxm, utm, prev(bnd["xhi"])[:, None, :]), uxp)
    uym = np.where(iy == 0, extrap_quadh(uyp, utm, prev(bnd["ylo"])[:, :, None]), uym)
    uyp = np.where(iy == ny - 1, extrap_quadh(uym, utm, prev(bnd["yhi"])[:, :, None]), uyp)
    v_new = (U - utm) / dt
    v_old = np.where(it == 1, init_ut[None], (utm - utmm) / dt)
    fu = (v_new - v_old) / dt - (uxm - 2 * utm + uxp) / dx ** 2 - (uym - 2 * utm + uyp) / dy ** 2
    u0 = init_u + 0.5 * dt * init_ut
    return np.where(it == 0, (U - u0[None]) * kimp, fu)


def loss_terms(values):
    """core.py:1093-1095: terms = mean(square(F_k)); loss = sum; norms = sqrt(terms)."""
    terms = [np.mean(np.square(v)) for v in values]
    return sum(terms), terms, [np.sqrt(t) for t in terms]


def numerical_jacobian_T(residual_fn, U, F):
    """A^T F for an affine residual, by explicit probing (small grids only; used to check plans)."""
    base = residual_fn(np.zeros_like(U))
    g = np.zeros_like(U)
    flat = g.reshape(-1)
    e = np.zeros_like(U)
    ef = e.reshape(-1)
    for j in range(U.size):
        ef[j] = 1
        flat[j] = np.sum((residual_fn(e) - base) * F)
        ef[j] = 0
    return g


# ------------------------------------------------------------------------------------------
# Optimizer updates
# ------------------------------------------------------------------------------------------
def adam_scalars(lr, beta_1, beta_2, t, dtype):
    """optimizer.py:307-314: constants are cast to `dtype` BEFORE the power / bias correction."""
    d = np.dtype(dtype).type
    lr, b1, b2, tt = d(lr), d(beta_1), d(beta_2), d(t)
    alpha = lr * np.sqrt(d(1) - b2 ** tt) / (d(1) - b1 ** tt)
    return d(alpha), d(d(1) - b1), d(d(1) - b2)


def adam_step(x, m, v, g, lr, t, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """optimizer.py:311-319."""
    dt = x.dtype
    alpha, omb1, omb2 = adam_scalars(lr, beta_1, beta_2, t, dt)
    m = m + (g - m) * omb1
    v = v + (np.square(g) - v) * omb2
    x = x - (m * alpha) / (np.sqrt(v) + dt.type(epsilon))
    return x, m, v


def gd_step(x, g, lr):
    """optimizer.py:269-270."""
    return x - g * x.dtype.type(lr)


# ------------------------------------------------------------------------------------------
# Whole evaluation: loss + gradient w.r.t. every multigrid term (core.py:1082-1104)
# ------------------------------------------------------------------------------------------
def eval_loss_grad_plan(terms, loc, offsets, table, rr, const, factors=None):
    U = mg_synthesize(terms, loc, factors) if len(terms) > 1 else terms[0] * (factors[0] if factors else 1)
    F = stencil_forward(U, offsets, table, rr, const)
    n = F.size
    loss = np.mean(np.square(F))
    gU = stencil_adjoint(F, offsets, table, rr, scale=2.0 / n)
    if len(terms) > 1:
        grads = mg_adjoint(gU, [t.shape for t in terms], loc, factors)
    else:
        grads = [gU * (factors[0] if factors else 1)]
    return loss, grads, F, U
