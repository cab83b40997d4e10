"""Host-side model of the z-slab decomposition of the z pass (numpy), driven by the correction
tables the C library exports (pbx_dist_tables_host).  Used by the CPU tests of the N>1 path: the
same moments / exchange / correction steps the CUDA path performs, on batches of z lines.

    out = Mzz c + Dzz d   on a periodic line of P * nzl points, rank p owning [p*nzl, (p+1)*nzl)
"""
import ctypes

import numpy as np

import poissbox_b200 as pbx
from poissbox_b200 import _lib

NB, RMAX = 48, 8


def composite_coef(kind, dz):
    """scheme constants as in pbx_coeffs.cu (src/compact_schemes.f90:188-190, 303-305)"""
    if kind == "D":
        a, b, al, s = 63.0 / 62.0 / dz, 17.0 / 62.0 / (3.0 * dz), 9.0 / 62.0, -1.0
    else:
        a, b, al, s = 0.75, 1.0 / 20.0, 0.3, 1.0
    r = (-1.0 + np.sqrt(1.0 - 4.0 * al * al)) / (2.0 * al)
    sc = (1.0 + r * r) ** 2
    c = [sc * s * (2 * a * a + 2 * b * b), sc * (a * a + 2 * s * a * b), sc * 2 * a * b, sc * b * b]
    if kind == "D":
        c[0] = -2.0 * (c[1] + c[2] + c[3])
    return c, r, al


def tables(nzl, dz):
    ncs, nrow = ctypes.c_int(), ctypes.c_int()
    R = (ctypes.c_int * 2)()
    arrs = [np.zeros((2, NB, RMAX)) for _ in range(5)]
    ptr = [a.ctypes.data_as(_lib._dp) for a in arrs]
    pbx.check(pbx.LIB.pbx_dist_tables_host(nzl, dz, ctypes.byref(ncs), ctypes.byref(nrow), R, *ptr))
    U, VnbM, VsM, VnbD, VsD = arrs
    return dict(ncs=ncs.value, nrow=nrow.value, R=(R[0], R[1]), U=U, VnbM=VnbM, VsM=VsM, VnbD=VnbD, VsD=VsD)


def _recursion(v, r):
    """causal double recursion from zero state along axis 0, then the anti-causal one"""
    out = np.array(v, dtype=np.float64)
    for sweep in (1, -1):
        y = np.zeros_like(out[0])
        z = np.zeros_like(out[0])
        idx = range(out.shape[0]) if sweep == 1 else range(out.shape[0] - 1, -1, -1)
        for i in idx:
            y = out[i] + r * y
            z = y + r * z
            out[i] = z
    return out


def _stencil_open(v, c):
    """7-point symmetric stencil along axis 0 with zero halos"""
    n = v.shape[0]
    pad = np.zeros((n + 6,) + v.shape[1:])
    pad[3:n + 3] = v
    out = c[0] * pad[3:n + 3]
    for k in (1, 2, 3):
        out = out + c[k] * (pad[3 - k:n + 3 - k] + pad[3 + k:n + 3 + k])
    return out


def local_open(c, d, dz):
    """what one rank computes on its slab alone: L_M c + L_D d (open line)"""
    cm, rm, _ = composite_coef("M", dz)
    cd, rd, _ = composite_coef("D", dz)
    return _stencil_open(_recursion(c, rm), cm) + _recursion(_stencil_open(d, cd), rd)


def moments(T, c, d):
    """(send_up, send_dn, self_a, self_b), each [RMAX, nlines]"""
    ncs = T["ncs"]
    up = T["VnbM"][0].T @ c[-NB:] + T["VnbD"][0].T @ d[-NB:]     # neighbour columns of the upper rank's block A
    dn = T["VnbM"][1].T @ c[:NB] + T["VnbD"][1].T @ d[:NB]       # ... of the lower rank's block B
    sa = T["VsM"][0][:ncs].T @ c[:ncs] + T["VsD"][0][:ncs].T @ d[:ncs]
    sb = T["VsM"][1][:ncs].T @ c[-ncs:] + T["VsD"][1][:ncs].T @ d[-ncs:]
    return up, dn, sa, sb


def correct(T, out, m_a, m_b):
    nrow = T["nrow"]
    out = out.copy()
    out[:nrow] += T["U"][0][:nrow] @ m_a
    out[-nrow:] += T["U"][1][:nrow] @ m_b
    return out


def periodic_truth(c, d, dz):
    """dense periodic evaluation of Mzz c + Dzz d"""
    n = c.shape[0]
    res = 0.0
    for kind, v in (("M", c), ("D", d)):
        cc, r, al = composite_coef(kind, dz)
        A = np.eye(n) + al * (np.roll(np.eye(n), 1, 0) + np.roll(np.eye(n), -1, 0))
        Ai2 = np.linalg.matrix_power(np.linalg.inv(A), 2) / (1.0 + r * r) ** 2

        def sten(w):
            return sum(cc[abs(k)] * np.roll(w, -k, 0) for k in range(-3, 4))

        res = res + (sten(Ai2 @ v) if kind == "M" else Ai2 @ sten(v))
    return res


# ------------------------------------------------------------------------------------------------
# The protocol the CUDA path uses (pbx_dist.cu: k_boundary, pbx_fast_common.cuh: zpass_body_slab):
# the neighbours' influence enters the slab's z pass through its native inputs -- recursion states
# and stencil halos -- rebuilt from nine numbers per z-line and direction.
# ------------------------------------------------------------------------------------------------
BM, BD = 48, 24


def _causal(u, r, Y=0.0, Z=0.0):
    y = np.zeros_like(u[0]) + Y
    z = np.zeros_like(u[0]) + Z
    out = np.empty_like(u)
    for i in range(u.shape[0]):
        y = u[i] + r * y
        z = y + r * z
        out[i] = z
    return out, y, z


def _anticausal(zv, r, W=0.0, X=0.0):
    w = np.zeros_like(zv[0]) + W
    x = np.zeros_like(zv[0]) + X
    out = np.empty_like(zv)
    for i in range(zv.shape[0] - 1, -1, -1):
        w = zv[i] + r * w
        x = w + r * x
        out[i] = x
    return out, w, x


def _stencil_halo(v, c, lo, hi):
    n = v.shape[0]
    pad = np.concatenate([lo, v, hi], 0)
    out = c[0] * pad[3:n + 3]
    for k in (1, 2, 3):
        out = out + c[k] * (pad[3 - k:n + 3 - k] + pad[3 + k:n + 3 + k])
    return out


def boundary_messages(c, d, dz):
    """(msg_up, msg_dn): nine arrays each, what k_boundary computes from one slab"""
    cm, rm, _ = composite_coef("M", dz)
    cd, rd, _ = composite_coef("D", dz)
    z3 = np.zeros((3,) + c.shape[1:])
    zM, yM, zMe = _causal(c[-BM:], rm)
    sD = _stencil_halo(d[-(BD + 3):], cd, z3, z3)[3:]
    _, yD, zD = _causal(sD, rd)
    up = [yM, zMe, zM[-2], zM[-3], yD, zD, d[-1], d[-2], d[-3]]
    j = np.arange(BM).reshape((-1,) + (1,) * (c.ndim - 1))
    A0M, A1M = np.sum(rm**j * c[:BM], 0), np.sum(j * rm**j * c[:BM], 0)
    sDb = _stencil_halo(d[:BD + 3], cd, z3, z3)[:BD]
    j = np.arange(BD).reshape((-1,) + (1,) * (c.ndim - 1))
    A0D, A1D = np.sum(rd**j * sDb, 0), np.sum(j * rd**j * sDb, 0)
    dn = [A0M, A1M, A0D, A1D, d[0], d[1], d[2], c[0], c[1]]
    return up, dn


def slab_zpass(c, d, dz, from_lo, from_up):
    """the slab's z pass with the neighbours' messages (zpass_body_slab)"""
    cm, rm, _ = composite_coef("M", dz)
    cd, rd, _ = composite_coef("D", dz)
    yMl, zMl, zMl1, zMl2, yDl, zDl, dl1, dl2, dl3 = from_lo
    A0M, A1M, A0D, A1D, du0, du1, du2, cu0, cu1 = from_up

    def G(a0, a1, r):
        q = r * r
        return a0 / (1 - q) ** 2, a0 * (1 + q) / (1 - q) ** 3 + a1 / (1 - q) ** 2

    def K(Y, Z, r):
        q = r * r
        return Z * r / (1 - q) + Y * r / (1 - q) ** 2, Z * r / (1 - q) ** 2 + Y * r * (1 + q) / (1 - q) ** 3

    s = _stencil_halo(d, cd, np.stack([dl3, dl2, dl1]), np.stack([du0, du1, du2]))
    s1, s2, s3 = cd[1] * d[0] + cd[2] * d[1] + cd[3] * d[2], cd[2] * d[0] + cd[3] * d[1], cd[3] * d[0]
    zDv, yDe, zDe = _causal(s, rd, yDl + s1 + rd * s2 + rd * rd * s3, zDl + s1 + 2 * rd * s2 + 3 * rd * rd * s3)
    e0, e1, e2 = cd[1] * d[-1] + cd[2] * d[-2] + cd[3] * d[-3], cd[2] * d[-1] + cd[3] * d[-2], cd[3] * d[-1]
    Wg, Xg = G(A0D + e0 + rd * e1 + rd * rd * e2, A1D + rd * e1 + 2 * rd * rd * e2, rd)
    Wk, Xk = K(yDe, zDe, rd)
    outD, _, _ = _anticausal(zDv, rd, Wg + Wk, Xg + Xk)

    zMv, yMe, zMe = _causal(c, rm, yMl, zMl)
    Wg, Xg = G(A0M, A1M, rm)
    Wk, Xk = K(yMe, zMe, rm)
    WT, XT = Wg + Wk, Xg + Xk
    xM, w, x = _anticausal(zMv, rm, WT, XT)
    lo = []
    for zt in (zMl, zMl1, zMl2):
        w = zt + rm * w
        x = w + rm * x
        lo.append(x)
    zt0 = cu0 + rm * (zMe + yMe)
    zt1 = cu1 + 2 * rm * cu0 + rm * rm * (zMe + 2 * yMe)
    w1, x1 = (WT - zt0) / rm, (XT - WT) / rm
    x2 = (x1 - w1) / rm
    del zt1
    return _stencil_halo(xM, cm, np.stack(lo[::-1]), np.stack([XT, x1, x2])) + outD
