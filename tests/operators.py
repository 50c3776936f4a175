"""
Problem operators used by the tests, written against the ODIL API exactly as a user would
(ctx.field / ctx.indices / mod.where / mod.roll).  They state the same discrete equations as the
reference examples (examples/poisson/poisson.py:57-123, examples/wave/wave.py:29-75).
"""
import argparse

import numpy as np

import odil


def dirichlet_neighbours(mod, u, um, up, idx, n, wall=0):
    """Ghost values across a wall at the face (half a cell away) by quadratic extrapolation."""
    ex = odil.core.extrap_quadh
    um2 = mod.where(idx == 0, ex(up, u, wall), um)
    up2 = mod.where(idx == n - 1, ex(um, u, wall), up)
    return um2, up2


def poisson_operator(ctx):
    mod, ndim = ctx.mod, ctx.domain.ndim
    h, idx, n = ctx.step(), ctx.indices(), ctx.size()
    if ndim == 1:
        h, idx, n = (h,) if np.ndim(h) == 0 else h, (idx,) if not isinstance(idx, tuple) else idx, n
    u = ctx.field("u")
    zero = mod.cast(0, u.dtype)
    lap = None
    for a in range(ndim):
        e = [1 if b == a else 0 for b in range(ndim)]
        um, up = ctx.field("u", *[-s for s in e]), ctx.field("u", *e)
        um, up = dirichlet_neighbours(mod, u, um, up, idx[a], n[a], zero)
        term = (up - 2 * u + um) / h[a] ** 2
        lap = term if lap is None else lap + term
    return [lap - ctx.extra.rhs]


def discrete_rhs(u, domain):
    """Applies the same discrete Laplacian to a known field (host side, through mod.roll)."""
    mod, ndim = domain.mod, domain.ndim
    h, idx, n = domain.step(), domain.indices(), domain.size()
    if ndim == 1:
        idx = (idx,) if not isinstance(idx, tuple) else idx
    zero = mod.cast(0, u.dtype)
    res = 0
    for a in range(ndim):
        um, up = mod.roll(u, 1, a), mod.roll(u, -1, a)
        um, up = dirichlet_neighbours(mod, u, um, up, idx[a], n[a], zero)
        res = res + (up - 2 * u + um) / h[a] ** 2
    return res


def hat_solution(domain):
    xs = domain.points()
    if domain.ndim == 1:
        xs = (xs,) if not isinstance(xs, tuple) else xs
    u = np.prod([(1 - np.asarray(x)) * np.asarray(x) * 5 for x in xs], axis=0)
    return (u ** 5 / (1 + u ** 5)) ** 0.2


def make_poisson(cshape, nlvl=0, dtype=np.float64, mg_nlvl=None):
    ndim = len(cshape)
    domain = odil.Domain(cshape=list(cshape), dimnames=["x", "y", "z", "w"][:ndim], multigrid=nlvl > 0,
                         mg_nlvl=nlvl if nlvl > 0 else None, dtype=dtype)
    ref_u = hat_solution(domain).astype(dtype)
    rhs = discrete_rhs(ref_u, domain)
    state = odil.State()
    state.fields["u"] = None
    state = domain.init_state(state)
    extra = argparse.Namespace(rhs=rhs, ref_u=ref_u)
    return odil.Problem(poisson_operator, domain, extra), state


def wave_exact(t, x):
    u, ut = 0, 0
    for i in range(1, 6):
        k = i * np.pi
        u = u + np.cos((x - t + 0.5) * k) + np.cos((x + t - 0.5) * k)
        ut = ut + k * np.sin((x - t + 0.5) * k) - k * np.sin((x + t - 0.5) * k)
    return u / 10, ut / 10


def wave_operator(ctx):
    """u_tt = u_xx on a (t, x) grid; Dirichlet data in x, initial u and u_t imposed in the first rows."""
    mod, extra = ctx.mod, ctx.extra
    dt, dx = ctx.step()
    it, ix = ctx.indices()
    nt, nx = ctx.size()
    u, utm, utmm = ctx.field("u"), ctx.field("u", -1, 0), ctx.field("u", -2, 0)
    uxm, uxp = ctx.field("u", -1, -1), ctx.field("u", -1, 1)
    left = mod.roll(extra.left_u, 1, axis=0)[:, None]
    right = mod.roll(extra.right_u, 1, axis=0)[:, None]
    ex = odil.core.extrap_quadh
    uxm = mod.where(ix == 0, ex(uxp, utm, left), uxm)
    uxp = mod.where(ix == nx - 1, ex(uxm, utm, right), uxp)
    v_new = (u - utm) / dt
    v_old = mod.where(it == 1, extra.init_ut[None, :], (utm - utmm) / dt)
    fu = (v_new - v_old) / dt - (uxm - 2 * utm + uxp) / dx ** 2
    u0 = extra.init_u + 0.5 * dt * extra.init_ut
    fu = mod.where(it == 0, (u - u0[None, :]) * extra.kimp, fu)
    return [("fu", fu)]


def make_wave(cshape, nlvl=0, dtype=np.float64):
    domain = odil.Domain(cshape=tuple(cshape), dimnames=("t", "x"), lower=(0, -1), upper=(1, 1), dtype=dtype,
                         multigrid=nlvl > 0, mg_nlvl=nlvl if nlvl > 0 else None)
    t1, x1 = domain.points_1d()
    left_u, _ = wave_exact(t1, t1 * 0 + domain.lower[1])
    right_u, _ = wave_exact(t1, t1 * 0 + domain.upper[1])
    init_u, init_ut = wave_exact(x1 * 0 + domain.lower[0], x1)
    extra = argparse.Namespace(kimp=1.0, left_u=left_u.astype(dtype), right_u=right_u.astype(dtype),
                               init_u=init_u.astype(dtype), init_ut=init_ut.astype(dtype))
    state = odil.State()
    state.fields["u"] = np.zeros(domain.cshape)
    state = domain.init_state(state)
    return odil.Problem(wave_operator, domain, extra), state


# --------------------------------------------------------------------------------------------------
# BASELINE configs[2]: the wave operator in two space dimensions, (t, x, y) grid
# --------------------------------------------------------------------------------------------------
def wave2_exact(t, x, y):
    """Sum of plane waves travelling along x, y and the diagonal: an exact solution of u_tt = u_xx + u_yy."""
    u, ut = 0, 0
    for i, (kx, ky) in enumerate([(1, 0), (0, 1), (1, 1), (2, -1), (1, 2)], start=1):
        w = np.pi * np.hypot(kx, ky)
        ph = np.pi * (kx * x + ky * y) + 0.3 * i
        u = u + np.cos(ph - w * t)
        ut = ut + w * np.sin(ph - w * t)
    return u / 10, ut / 10


def wave2_operator(ctx):
    """u_tt = u_xx + u_yy on a (t, x, y) grid, written with the calls of examples/wave/wave.py:29-75."""
    mod, extra = ctx.mod, ctx.extra
    dt, dx, dy = ctx.step()
    it, ix, iy = ctx.indices()
    nt, nx, ny = ctx.size()
    u, utm, utmm = ctx.field("u"), ctx.field("u", -1, 0, 0), ctx.field("u", -2, 0, 0)
    uxm, uxp = ctx.field("u", -1, -1, 0), ctx.field("u", -1, 1, 0)
    uym, uyp = ctx.field("u", -1, 0, -1), ctx.field("u", -1, 0, 1)
    ex = odil.core.extrap_quadh

    def prev(a):  # boundary data at time level t-1
        return mod.roll(a, 1, axis=0)

    uxm = mod.where(ix == 0, ex(uxp, utm, prev(extra.xlo)[:, None, :]), uxm)
    uxp = mod.where(ix == nx - 1, ex(uxm, utm, prev(extra.xhi)[:, None, :]), uxp)
    uym = mod.where(iy == 0, ex(uyp, utm, prev(extra.ylo)[:, :, None]), uym)
    uyp = mod.where(iy == ny - 1, ex(uym, utm, prev(extra.yhi)[:, :, None]), uyp)
    v_new = (u - utm) / dt
    v_old = mod.where(it == 1, extra.init_ut[None], (utm - utmm) / dt)
    fu = (v_new - v_old) / dt - (uxm - 2 * utm + uxp) / dx ** 2 - (uym - 2 * utm + uyp) / dy ** 2
    u0 = extra.init_u + 0.5 * dt * extra.init_ut
    fu = mod.where(it == 0, (u - u0[None]) * extra.kimp, fu)
    return [("fu", fu)]


def make_wave2(cshape, dtype=np.float64):
    domain = odil.Domain(cshape=tuple(cshape), dimnames=("t", "x", "y"), lower=(0, -1, -1), upper=(1, 1, 1),
                         dtype=dtype, multigrid=False)
    t1, x1, y1 = domain.points_1d()
    T, X, Y = np.meshgrid(t1, x1, y1, indexing="ij")
    lo, hi = domain.lower, domain.upper
    extra = argparse.Namespace(kimp=1.0)
    extra.xlo = wave2_exact(T[:, 0, :], lo[1], Y[:, 0, :])[0].astype(dtype)
    extra.xhi = wave2_exact(T[:, 0, :], hi[1], Y[:, 0, :])[0].astype(dtype)
    extra.ylo = wave2_exact(T[:, :, 0], X[:, :, 0], lo[2])[0].astype(dtype)
    extra.yhi = wave2_exact(T[:, :, 0], X[:, :, 0], hi[2])[0].astype(dtype)
    u0, ut0 = wave2_exact(lo[0], X[0], Y[0])
    extra.init_u, extra.init_ut = u0.astype(dtype), ut0.astype(dtype)
    extra.ref_u = wave2_exact(T, X, Y)[0]
    state = odil.State()
    state.fields["u"] = np.zeros(domain.cshape)
    state = domain.init_state(state)
    return odil.Problem(wave2_operator, domain, extra), state
