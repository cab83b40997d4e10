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
