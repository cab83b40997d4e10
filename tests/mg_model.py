"""numpy model of the multigrid preconditioner of poissbox_b200/csrc/pbx_mg.cu (TEST
INFRASTRUCTURE): V(nu, nu) cycle on S = -(2nd-order star), damped Jacobi (omega = 6/7),
cell-centred trilinear prolongation, restriction = its transpose / 8, re-discretised coarse
operators, 30 Jacobi sweeps on the coarsest grid; and the preconditioned CG loop with PETSc's
KSPCG semantics (preconditioned norm, constant null space removed from z)."""
import numpy as np

OMEGA = 6.0 / 7.0
COARSE_SWEEPS = 30


def s_apply(z, h):
    out = np.zeros_like(z)
    for ax in range(3):
        out += (1.0 / h[ax] ** 2) * (2 * z - np.roll(z, 1, ax) - np.roll(z, -1, ax))
    return out


def restrict(r):
    for ax in range(3):
        n = r.shape[ax]
        a = np.take(r, np.arange(0, n, 2), axis=ax)      # fine 2I
        b = np.take(r, np.arange(1, n, 2), axis=ax)      # fine 2I+1
        r = 0.125 * np.roll(b, 1, axis=ax) + 0.375 * a + 0.375 * b + 0.125 * np.roll(a, -1, axis=ax)
    return r


def prolong(c):
    for ax in range(3):
        lo = 0.75 * c + 0.25 * np.roll(c, 1, ax)
        hi = 0.75 * c + 0.25 * np.roll(c, -1, ax)
        shp = list(c.shape)
        shp[ax] *= 2
        f = np.empty(shp)
        sl = [slice(None)] * 3
        sl[ax] = slice(0, None, 2)
        f[tuple(sl)] = lo
        sl[ax] = slice(1, None, 2)
        f[tuple(sl)] = hi
        c = f
    return c


def vcycle(r, h, nu=2):
    n = r.shape
    wd = OMEGA / (2 * sum(1.0 / hh**2 for hh in h))
    if min(n) <= 4 or any(m % 2 for m in n):
        z = wd * r
        for _ in range(COARSE_SWEEPS - 1):
            z = z + wd * (r - s_apply(z, h))
        return z
    z = wd * r
    for _ in range(nu - 1):
        z = z + wd * (r - s_apply(z, h))
    z = z + prolong(vcycle(restrict(r - s_apply(z, h)), tuple(2 * hh for hh in h), nu))
    for _ in range(nu):
        z = z + wd * (r - s_apply(z, h))
    return z


def pc_apply(r, h, nu=2):
    z = vcycle(r - r.mean(), h, nu)
    return z - z.mean()


def pcg(apply_a, b, pc, rtol=1e-8, maxit=10000):
    """returns x, its, history of ||z||"""
    x = np.zeros_like(b)
    r = b.copy()
    z = pc(r)
    dp0 = np.linalg.norm(z)
    beta = np.vdot(z, r)
    p = z.copy()
    hist = [dp0]
    for it in range(1, maxit + 1):
        w = apply_a(p)
        a = beta / np.vdot(p, w)
        x += a * p
        r -= a * w
        z = pc(r)
        dp = np.linalg.norm(z)
        hist.append(dp)
        if dp <= rtol * dp0:
            return x, it, np.array(hist)
        bn = np.vdot(z, r)
        p = z + (bn / beta) * p
        beta = bn
    return x, maxit, np.array(hist)
