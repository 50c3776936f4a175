"""
NumPy restatements of the INDEX LOGIC of the shared-memory tile kernels (odil_b200/csrc/tile2d.cuh fused mode,
mg_tile2d.cuh synthesis and transpose, tile3d.cuh marching rings): same tile sizes, halo widths, staging order, magic-number row split,
wrap / clamp / reflect rules and closed-form pad fold, executed tile by tile on the host and compared with the
oracle.  This is how the kernels' addressing was checked before they first ran on a GPU (they then passed their
parity tests unchanged); it stays as a guard on the design -- the kernels themselves are tested in
tests/test_kernels_gpu.py.
"""
import numpy as np
import pytest

from oracle import odil_oracle as orc

TY, TX = 32, 64      # kT2Y, kT2X
CY, CX = 16, 64      # kM2Y, kM2X


def _wrap(i, n):     # t2_wrap: C remainder, then shift into range
    if i < 0 or i >= n:
        i = int(np.fmod(i, n))
        if i < 0:
            i += n
    return i


def _cls(i, n, r):   # t2_class
    if i < r:
        return i
    d = n - 1 - i
    return 2 * r - d if d < r else r


def tile2d_fused(U, c, table, offs, R, scale):
    N0, N1 = U.shape
    noff = len(offs)
    H0, H1 = max(abs(o[0]) for o in offs), max(abs(o[1]) for o in offs)
    AH, AW, FH, FW = TY + 4 * H0, TX + 4 * H1, TY + 2 * H0, TX + 2 * H1
    mA, mF = (1 << 32) // AW + 1, (1 << 32) // FW + 1
    C1 = 2 * R[1] + 1
    tab = table.reshape(-1, noff)
    G, Fo, ss = np.full(U.shape, np.nan), np.full(U.shape, np.nan), 0.0
    for ty0 in range(0, N0, TY):
        for tx0 in range(0, N1, TX):
            sA, sF, sC = np.zeros(AH * AW), np.zeros(FH * FW), np.zeros(FH * FW, dtype=int)
            sDA = [o[0] * AW + o[1] for o in offs]
            sDF = [o[0] * FW + o[1] for o in offs]
            for e in range(AH * AW):
                r = (e * mA) >> 32
                assert r == e // AW
                sA[e] = U[_wrap(ty0 - 2 * H0 + r, N0), _wrap(tx0 - 2 * H1 + e - r * AW, N1)]
            for e in range(FH * FW):
                r = (e * mF) >> 32
                cc = e - r * FW
                ly, lx = ty0 - H0 + r, tx0 - H1 + cc
                gy, gx = _wrap(ly, N0), _wrap(lx, N1)
                cl = _cls(gy, N0, R[0]) * C1 + _cls(gx, N1, R[1])
                at = (r + H0) * AW + cc + H1
                f = c[gy, gx] + sum(tab[cl, o] * sA[at + sDA[o]] for o in range(noff))
                sF[e], sC[e] = f, cl
                if H0 <= r < H0 + TY and H1 <= cc < H1 + TX and ly < N0 and lx < N1:
                    ss += f * f
                    Fo[ly, lx] = f
            for e in range(TY * TX):
                r, cc = divmod(e, TX)
                y, x = ty0 + r, tx0 + cc
                if y < N0 and x < N1:
                    at = (r + H0) * FW + cc + H1
                    G[y, x] = scale * sum(tab[sC[at - sDF[o]], o] * sF[at - sDF[o]] for o in range(noff))
    return Fo, G, ss


@pytest.mark.parametrize("shape,offs,R", [
    ((16, 12), [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)], (2, 1)),
    ((33, 70), [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], (2, 1)),
    ((3, 5), [(0, 0), (2, -2), (-1, 1)], (1, 2)),
    ((40, 66), [(0, 0), (3, 0), (0, -4), (-2, 2)], (0, 0)),
])
def test_tile2d_fused_addressing(shape, offs, R):
    rng = np.random.default_rng(0)
    table = rng.standard_normal(tuple(2 * r + 1 for r in R) + (len(offs),))
    U, c = rng.standard_normal(shape), rng.standard_normal(shape)
    F_ref = orc.stencil_forward(U, offs, table, R, c)
    g_ref = orc.stencil_adjoint(F_ref, offs, table, R, 0.37)
    Fo, G, ss = tile2d_fused(U, c, table, offs, R, 0.37)
    assert np.abs(Fo - F_ref).max() < 1e-12 and np.abs(G - g_ref).max() < 1e-12
    assert abs(ss - (F_ref ** 2).sum()) < 1e-9 * (F_ref ** 2).sum()


def _clamp(q, n):
    return 0 if q < 0 else (n - 1 if q > n - 1 else q)


def _reflect(q, n):
    return 1 if q < 0 else (n - 2 if q > n - 1 else q)


def interp_add2t(coarse, term, cfac, ffac):
    n0, n1 = coarse.shape
    PW = CX + 2
    out = np.full((2 * n0, 2 * n1), np.nan)
    for cy0 in range(0, n0, CY):
        for cx0 in range(0, n1, CX):
            sP = np.zeros((CY + 2) * PW)
            for e in range((CY + 2) * PW):
                r, cc = divmod(e, PW)
                qy, qx = min(cy0 - 1 + r, n0), min(cx0 - 1 + cc, n1)
                sy, sx = _clamp(qy, n0), _clamp(qx, n1)
                v = coarse[sy, sx]
                if sy != qy or sx != qx:
                    v = 2 * v - coarse[_reflect(qy, n0), _reflect(qx, n1)]
                sP[e] = v
            for vec in range(2 * CY * (CX // 2)):
                fr, vx = divmod(vec, CX // 2)
                I, a = fr >> 1, fr & 1
                fy, lc = 2 * cy0 + fr, 2 * vx
                fx = 2 * (cx0 + lc)
                if fy >= 2 * n0 or fx >= 2 * n1:
                    continue
                near, far = (I + 1) * PW + lc, (I + 2 * a) * PW + lc
                h = [3 * sP[near + k] + sP[far + k] for k in range(4)]
                o = [3 * h[1] + h[0], 3 * h[1] + h[2], 3 * h[2] + h[1], 3 * h[2] + h[3]]
                for k in range(4):
                    out[fy, fx + k] = cfac * (o[k] / 16) + ffac * term[fy, fx + k]
    return out


def interp_adjoint2t(gf, n0, n1, scale):
    FH, FW, GW = 2 * (CY + 4) + 2, 2 * (CX + 4) + 2, CX + 4
    g = np.full((n0, n1), np.nan)
    for cy0 in range(0, n0, CY):
        for cx0 in range(0, n1, CX):
            fy0, fx0 = 2 * (cy0 - 2) - 1, 2 * (cx0 - 2) - 1
            sG = np.zeros(FH * FW)
            for e in range(FH * FW):
                r, cc = divmod(e, FW)
                y, x = fy0 + r, fx0 + cc
                if 0 <= y < 2 * n0 and 0 <= x < 2 * n1:
                    sG[e] = gf[y, x]
            gp = np.zeros((CY + 4) * GW)
            w = (1, 3, 3, 1)
            for e in range((CY + 4) * GW):
                r, cc = divmod(e, GW)
                qy, qx = cy0 - 2 + r, cx0 - 2 + cc
                if -1 <= qy <= n0 and -1 <= qx <= n1:
                    gp[e] = sum(w[i] * w[j] * sG[(2 * r + i) * FW + 2 * cc + j] for i in range(4) for j in range(4)) / 16
            for e in range(CY * CX):
                r, cc = divmod(e, CX)
                cy, cx = cy0 + r, cx0 + cc
                if cy >= n0 or cx >= n1:
                    continue
                b = (r + 2) * GW + cc + 2

                def cl(at):
                    return gp[at] + (gp[at - 1] if cx == 0 else 0) + (gp[at + 1] if cx == n1 - 1 else 0)

                def rf(at):
                    return gp[at] + (gp[at - 2] if cx == 1 else 0) + (gp[at + 2] if cx == n1 - 2 else 0)

                scl = cl(b) + (cl(b - GW) if cy == 0 else 0) + (cl(b + GW) if cy == n0 - 1 else 0)
                srf = rf(b) + (rf(b - 2 * GW) if cy == 1 else 0) + (rf(b + 2 * GW) if cy == n0 - 2 else 0)
                g[cy, cx] = scale * (2 * scl - srf)
    return g


@pytest.mark.parametrize("cshape", [(2, 2), (3, 4), (17, 66), (5, 130)])
def test_mg_tile2d_addressing_and_pad_fold(cshape):
    rng = np.random.default_rng(1)
    n0, n1 = cshape
    u, t = rng.standard_normal(cshape), rng.standard_normal((2 * n0, 2 * n1))
    ref = 0.7 * orc.interp_to_finer(u, "cc") + 1.3 * t
    assert np.abs(interp_add2t(u, t, 0.7, 1.3) - ref).max() < 1e-13
    aref = 0.9 * orc.interp_adjoint(t, "cc", cshape)
    assert np.abs(interp_adjoint2t(t, n0, n1, 0.9) - aref).max() < 1e-13


# --------------------------------------------------------------------------------------------------
# tile3d.cuh: 16 x 64 tiles marching along axis 0 with rings of 2*H0+1 U and F planes
# --------------------------------------------------------------------------------------------------
def _slot(p, nr):    # t3_slot
    p = int(np.fmod(p, nr))
    return p + nr if p < 0 else p


def tile3d_fused(U, c, table, offs, R, scale, zchunk):
    T3Y, T3X = 16, 64
    N0, N1, N2 = U.shape
    noff = len(offs)
    H0, H1, H2 = (max(abs(o[a]) for o in offs) for a in range(3))
    AH, AW, FH, FW, NR = T3Y + 4 * H1, T3X + 4 * H2, T3Y + 2 * H1, T3X + 2 * H2, 2 * H0 + 1
    C1, C2 = 2 * R[1] + 1, 2 * R[2] + 1
    tab = table.reshape(-1, noff)
    G, Fo, ss = np.full(U.shape, np.nan), np.full(U.shape, np.nan), 0.0
    for zs in range(0, N0, zchunk):
        ze = min(zs + zchunk, N0)
        for ty0 in range(0, N1, T3Y):
            for tx0 in range(0, N2, T3X):
                sU, sF = np.zeros((NR, AH * AW)), np.zeros((NR, FH * FW))
                sC = np.zeros((NR, FH * FW), dtype=int)

                def stage(pz):
                    rows = [_wrap(ty0 - 2 * H1 + r, N1) for r in range(AH)]
                    cols = [_wrap(tx0 - 2 * H2 + q, N2) for q in range(AW)]
                    sU[_slot(pz, NR)] = U[_wrap(pz, N0)][np.ix_(rows, cols)].reshape(-1)

                j0, j1 = zs - H0, ze - 1 + H0
                for pz in range(j0 - H0, j0 + H0):
                    stage(pz)
                for j in range(j0, j1 + 1):
                    k = j - H0
                    stage(j + H0)
                    oU = [(_slot(j + o[0], NR), o[1] * AW + o[2]) for o in offs]
                    oF = [(_slot(k - o[0], NR), -(o[1] * FW + o[2])) for o in offs]
                    gz, fs = _wrap(j, N0), _slot(j, NR)
                    for e in range(FH * FW):
                        r, cc = divmod(e, FW)
                        ly, lx = ty0 - H1 + r, tx0 - H2 + cc
                        gy, gx = _wrap(ly, N1), _wrap(lx, N2)
                        cl = (_cls(gz, N0, R[0]) * C1 + _cls(gy, N1, R[1])) * C2 + _cls(gx, N2, R[2])
                        at = (r + H1) * AW + cc + H2
                        f = c[gz, gy, gx] + sum(tab[cl, o] * sU[oU[o][0], at + oU[o][1]] for o in range(noff))
                        sF[fs, e], sC[fs, e] = f, cl
                        if zs <= j < ze and H1 <= r < H1 + T3Y and H2 <= cc < H2 + T3X and ly < N1 and lx < N2:
                            ss += f * f
                            Fo[j, ly, lx] = f
                    if k >= zs:
                        for e in range(T3Y * T3X):
                            r, cc = divmod(e, T3X)
                            y, x = ty0 + r, tx0 + cc
                            if y < N1 and x < N2:
                                at = (r + H1) * FW + cc + H2
                                G[k, y, x] = scale * sum(tab[sC[s, at + dd], o] * sF[s, at + dd]
                                                         for o, (s, dd) in enumerate(oF))
    return Fo, G, ss


WAVE2 = [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)]


@pytest.mark.parametrize("shape,offs,R,zchunk", [
    ((7, 10, 9), WAVE2, (2, 1, 1), 3),
    ((5, 18, 70), WAVE2, (2, 1, 1), 64),
    ((4, 5, 6), [(0, 0, 0), (1, 1, 1), (-2, 0, 2), (0, -1, 0)], (1, 1, 2), 2),   # planes wrap more than once
    ((3, 4, 5), [(0, 0, 0), (0, 1, -1)], (0, 0, 0), 1),                           # no coupling along axis 0
])
def test_tile3d_ring_addressing(shape, offs, R, zchunk):
    rng = np.random.default_rng(3)
    table = rng.standard_normal(tuple(2 * r + 1 for r in R) + (len(offs),))
    U, c = rng.standard_normal(shape), rng.standard_normal(shape)
    F_ref = orc.stencil_forward(U, offs, table, R, c)
    g_ref = orc.stencil_adjoint(F_ref, offs, table, R, 0.37)
    Fo, G, ss = tile3d_fused(U, c, table, offs, R, 0.37, zchunk)
    assert np.abs(Fo - F_ref).max() < 1e-12 and np.abs(G - g_ref).max() < 1e-12
    assert abs(ss - (F_ref ** 2).sum()) < 1e-9 * (F_ref ** 2).sum()


# --------------------------------------------------------------------------------------------------------------------
# k_tile2w (tile2w.cuh): one warp per (strip of 120 columns, chunk of rows), marching down the rows with two private
# rings stored cell-major; lanes 0 and 31 carry the halo columns.  The emulation runs warp by warp with the lanes as a
# vector of 32, the rings as arrays [8 slots][4 cell planes][34 words], the byte-offset tables kU / kF of the launcher
# (in words here), the two fast paths and the table paths.
# --------------------------------------------------------------------------------------------------------------------
W2_VW, W2_OWN, W2_W = 4, 120, 34


def _fdiv4(v):
    return v // 4  # Python floors; the launcher spells it out for C++


def wrap_free_table(table, offs, R):
    """Zeroes every (class, offset) entry whose neighbour would cross the boundary -- the rule of plan_create
    (stencil.cu: wrap_free)."""
    t = table.copy()
    nd = len(R)
    for cl in np.ndindex(*t.shape[:-1]):
        for o, off in enumerate(offs):
            for a in range(nd):
                d, r, c = off[a], R[a], cl[a]
                if d == 0:
                    continue
                crosses = (c + d < 0) if c < r else ((2 * r - c) < d if c > r else abs(d) > r)
                if crosses:
                    t[cl + (o,)] = 0.0
    return t


def tile2w_fused(U, c, table, offs, R, scale, rows_per_chunk):
    N0, N1 = U.shape
    assert N1 % 4 == 0
    noff = len(offs)
    H0, H1 = max(abs(o[0]) for o in offs), max(abs(o[1]) for o in offs)
    assert H0 <= 2 and H1 <= 2 and noff <= 8
    NR = 2 * H0 + 1                       # rows per ring
    SLOT = 4 * W2_W
    C1 = 2 * R[1] + 1
    tab = table.reshape(-1, noff)
    CI = R[0] * C1 + R[1]
    kU = [[((j + o[1]) - 4 * _fdiv4(j + o[1])) * W2_W + _fdiv4(j + o[1]) for j in range(4)] for o in offs]
    kF = [[((j - o[1]) - 4 * _fdiv4(j - o[1])) * W2_W + _fdiv4(j - o[1]) for j in range(4)] for o in offs]
    rU = [o[0] - H0 for o in offs]        # ring slot relative to the newest row
    rF = [-H0 - o[0] for o in offs]
    nstrips = (N1 + W2_OWN - 1) // W2_OWN
    nchunks = (N0 + rows_per_chunk - 1) // rows_per_chunk
    G, Fo, ss = np.full(U.shape, np.nan), np.full(U.shape, np.nan), 0.0
    lanes = np.arange(32)
    for item in range(nstrips * nchunks):
        strip, chunk = item % nstrips, item // nstrips
        ys, ye = chunk * rows_per_chunk, min((chunk + 1) * rows_per_chunk, N0)
        x0 = strip * W2_OWN + W2_VW * (lanes - 1)
        xin = (x0 >= 0) & (x0 < N1)
        own = xin & (lanes >= 1) & (lanes <= 30)
        ccls = np.array([[(_cls(x0[l] + j, N1, R[1]) if xin[l] else 0) for j in range(4)] for l in range(32)])
        fastXF = bool(np.all(~xin | ((x0 >= R[1]) & (x0 + 3 < N1 - R[1]))))
        fastXG = bool(np.all(~own | ((x0 >= R[1] + H1) & (x0 + 3 < N1 - R[1] - H1))))
        ringU = np.zeros(NR * SLOT)
        ringF = np.zeros(NR * SLOT)
        base = lanes + 1  # word of (slot 0, plane 0, own lane)

        def load(A, r, extra=True):
            out = np.zeros((32, 4))
            if 0 <= r < N0 and extra:
                for l in range(32):
                    if xin[l]:
                        out[l] = A[r, x0[l]: x0[l] + 4]
            return out

        def stage(u, slot):
            for j in range(4):
                ringU[slot * SLOT + j * W2_W + base] = u[:, j]

        def slot_of(cu, rel):
            t = cu + rel
            assert -NR <= t < NR
            return t + NR if t < 0 else t

        for q in range(2 * H0):           # prologue
            stage(load(U, ys - 2 * H0 + q), q)
        cu = NR - 1
        for ru in range(ys, ye + 2 * H0):
            stage(load(U, ru), cu)
            jf = ru - H0
            f = load(c, jf, jf < ye + H0)
            rin = 0 <= jf < N0
            if rin:
                so = [slot_of(cu, r) * SLOT for r in rU]
                if fastXF and R[0] <= jf < N0 - R[0]:
                    for o in range(noff):
                        for j in range(4):
                            f[:, j] += tab[CI, o] * ringU[so[o] + kU[o][j] + base]
                else:
                    rc = _cls(jf, N0, R[0]) * C1
                    for o in range(noff):
                        for j in range(4):
                            f[:, j] += tab[rc + ccls[:, j], o] * ringU[so[o] + kU[o][j] + base]
            if not rin:
                f[:] = 0
            f[~xin] = 0
            if ys <= jf < ye:
                ss += float((f[own] ** 2).sum())
                for l in lanes[own]:
                    Fo[jf, x0[l]: x0[l] + 4] = f[l]
            for j in range(4):
                ringF[cu * SLOT + j * W2_W + base] = f[:, j]
            k = ru - 2 * H0
            if k >= ys:
                g = np.zeros((32, 4))
                so = [slot_of(cu, r) * SLOT for r in rF]
                if fastXG and R[0] + H0 <= k < N0 - R[0] - H0:
                    for o in range(noff):
                        for j in range(4):
                            g[:, j] += tab[CI, o] * ringF[so[o] + kF[o][j] + base]
                else:
                    for o, off in enumerate(offs):
                        sy = k - off[0]
                        if not 0 <= sy < N0:
                            continue
                        rc = _cls(sy, N0, R[0]) * C1
                        for j in range(4):
                            sx = x0 + j - off[1]
                            cx = np.array([(_cls(v, N1, R[1]) if 0 <= v < N1 else 0) for v in sx])
                            g[:, j] += tab[rc + cx, o] * ringF[so[o] + kF[o][j] + base]
                for l in lanes[own]:
                    G[k, x0[l]: x0[l] + 4] = g[l] * scale
            cu = 0 if cu + 1 == NR else cu + 1
    return Fo, G, ss


@pytest.mark.parametrize("shape,offs,R,rpc", [
    ((16, 12), [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)], (2, 1), 8),      # wave footprint (wave.py:38-46)
    ((33, 124), [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], (1, 1), 8),        # 5-point star, two strips
    ((9, 248), [(0, 0), (2, -2), (-1, 1), (1, 2), (-2, 0), (0, -1)], (2, 2), 16),
    ((40, 120), [(0, 0), (1, 0), (0, -2), (-2, 2), (2, 1), (-1, -1), (0, 1), (1, -2)], (3, 2), 24),
    ((5, 4), [(0, 0), (0, 1), (1, 0)], (1, 1), 8),
    ((24, 368), [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)], (0, 0), 8),         # no boundary classes at all
])
def test_tile2w_fused_addressing(shape, offs, R, rpc):
    rng = np.random.default_rng(1)
    table = wrap_free_table(rng.standard_normal(tuple(2 * r + 1 for r in R) + (len(offs),)), offs, R)
    if R == (0, 0):  # every cell has the interior class: only offsets that never leave the array are wrap-free
        offs, table = [(0, 0)], table[..., :1]
    U, c = rng.standard_normal(shape), rng.standard_normal(shape)
    F_ref = orc.stencil_forward(U, offs, table, R, c)
    g_ref = orc.stencil_adjoint(F_ref, offs, table, R, 0.37)
    Fo, G, ss = tile2w_fused(U, c, table, offs, R, 0.37, rpc)
    assert np.abs(Fo - F_ref).max() < 1e-12 and np.abs(G - g_ref).max() < 1e-12
    assert abs(ss - (F_ref ** 2).sum()) < 1e-9 * (F_ref ** 2).sum()
