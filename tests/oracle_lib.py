"""ctypes binding of the CPU oracle (oracle/pbx_oracle.c) -- test infrastructure only.

Arrays are numpy float64 in Fortran (column-major) order: f[i, j, k] with i fastest, the layout of
the reference's `f(i,j,k)` (src/compact_schemes.f90:19-23).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ODIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ODIR, "_build", "libpbx_oracle.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _build():
    src = os.path.join(_ODIR, "pbx_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _ODIR])


def _load():
    _build()
    lib = ctypes.CDLL(_SO)
    i, d = ctypes.c_int, ctypes.c_double
    sigs = {
        "orc_fwd_sweep": (None, [i, _dp, _dp, _dp, _dp]),
        "orc_bwd_sweep": (None, [i, _dp, _dp, _dp]),
        "orc_tdma": (None, [i, _dp, _dp, _dp, _dp]),
        "orc_tdma_periodic": (None, [i, _dp, _dp, _dp, _dp]),
        "orc_eval_1d_rhs": (None, [d, d, i, i, i, _dp, _dp]),
        "orc_grad_1d": (i, [i, _dp, d, i, _dp, i]),
        "orc_div_1d": (i, [i, _dp, d, i, _dp]),
        "orc_interp_1d": (i, [i, _dp, i, _dp, i]),
        "orc_interp_1d_div": (i, [i, _dp, i, _dp]),
        "orc_grad": (None, [i, i, i, _dp, _dp, _dp]),
        "orc_div": (None, [i, i, i, _dp, _dp, _dp]),
        "orc_interp": (None, [i, i, i, _dp, _dp, i]),
        "orc_interp_div": (None, [i, i, i, _dp, _dp]),
        "orc_lapl": (None, [i, i, i, _dp, _dp, _dp]),
        "orc_lapl_1d_coeffs": (None, [d, _dp]),
        "orc_lapl_star_coeffs": (None, [d, d, d, _dp]),
        "orc_evaluate_laplacian_pointwise": (d, [_dp, _dp]),
        "orc_star": (None, [i, i, i, _dp, _dp, _dp]),
        "orc_set_threads": (None, [i]),
        "orc_get_threads": (i, []),
        "orc_cg_solve": (i, [i, i, i, _dp, _dp, _dp, d, d, i, _dp, _ip, _dp, i]),
        "orc_cg_solve_op": (i, [i, i, i, i, _dp, _dp, _dp, d, d, i, _dp, _ip, _dp, i]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


LIB = _load()


def _p(a):
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def set_threads(n):
    LIB.orc_set_threads(int(n))


# ---- tridsol ----
def fwd_sweep(a, b, c, d):
    a, b, c, d = (np.array(v, dtype=np.float64) for v in (a, b, c, d))
    LIB.orc_fwd_sweep(len(d), _p(a), _p(b), _p(c), _p(d))
    return b, d


def bwd_sweep(b, c, d):
    b, c, d = (np.array(v, dtype=np.float64) for v in (b, c, d))
    LIB.orc_bwd_sweep(len(d), _p(b), _p(c), _p(d))
    return d


def tdma(a, b, c, d):
    a, b, c, d = (np.array(v, dtype=np.float64) for v in (a, b, c, d))
    LIB.orc_tdma(len(d), _p(a), _p(b), _p(c), _p(d))
    return d


def tdma_periodic(a, b, c, d):
    a, b, c, d = (np.array(v, dtype=np.float64) for v in (a, b, c, d))
    LIB.orc_tdma_periodic(len(d), _p(a), _p(b), _p(c), _p(d))
    return d


# ---- compact_schemes 1-D ----
def eval_1d_rhs(a, b, opsign, stagger, f):
    f = np.array(f, dtype=np.float64)
    rhs = np.full_like(f, 73.29)
    LIB.orc_eval_1d_rhs(a, b, opsign, stagger, len(f), _p(f), _p(rhs))
    return rhs


def grad_1d(f, dx, stagger=-1):
    f = np.array(f, dtype=np.float64)
    df = np.full_like(f, 73.29)
    assert LIB.orc_grad_1d(len(f), _p(f), dx, len(df), _p(df), stagger) == 0
    return df


def div_1d(f, dx):
    return grad_1d(f, dx, +1)


def interp_1d(f, stagger=-1):
    f = np.array(f, dtype=np.float64)
    fi = np.full_like(f, 73.29)
    assert LIB.orc_interp_1d(len(f), _p(f), len(fi), _p(fi), stagger) == 0
    return fi


def interp_1d_div(f):
    return interp_1d(f, +1)


# ---- compact_schemes 3-D ----
def _dx(dx):
    return (ctypes.c_double * 3)(*[float(v) for v in dx])


def grad(f, dx):
    f = _f(f)
    nx, ny, nz = f.shape
    df = np.full((nx, ny, nz, 3), 73.29, order="F")
    LIB.orc_grad(nx, ny, nz, _p(f), _dx(dx), _p(df))
    return df


def div(f, dx):
    f = _f(f)
    nx, ny, nz, _ = f.shape
    df = np.full((nx, ny, nz), 73.29, order="F")
    LIB.orc_div(nx, ny, nz, _p(f), _dx(dx), _p(df))
    return df


def interp(f, stagger=-1):
    f = _f(f)
    nx, ny, nz = f.shape
    fi = np.full((nx, ny, nz), 73.29, order="F")
    LIB.orc_interp(nx, ny, nz, _p(f), _p(fi), stagger)
    return fi


def interp_div(f):
    return interp(f, +1)


def lapl(f, dx):
    f = _f(f)
    nx, ny, nz = f.shape
    out = np.full((nx, ny, nz), 73.29, order="F")
    LIB.orc_lapl(nx, ny, nz, _p(f), _dx(dx), _p(out))
    return out


# ---- the 2nd-order star (coefficients.f90, compute_lapl) ----
def lapl_1d_coeffs(dx):
    c = np.zeros(3)
    LIB.orc_lapl_1d_coeffs(float(dx), _p(c))
    return c


def lapl_star_coeffs(dx, dy, dz):
    c = np.zeros((3, 3, 3), order="F")
    LIB.orc_lapl_star_coeffs(float(dx), float(dy), float(dz), _p(c))
    return c


def evaluate_laplacian_pointwise(f, grid_deltas):
    f = _f(f)
    assert f.shape == (3, 3, 3)
    return LIB.orc_evaluate_laplacian_pointwise(_p(f), _dx(grid_deltas))


def star(x, dx):
    """compute_lapl_pointwise on a periodic box (what mfmult applies today)"""
    x = _f(x)
    nx, ny, nz = x.shape
    out = np.full((nx, ny, nz), 73.29, order="F")
    LIB.orc_star(nx, ny, nz, _p(x), _dx(dx), _p(out))
    return out


def cg_solve(b, dx, rtol=1e-5, abstol=1e-50, maxit=10000, op=0):
    """op = 0: compact Laplacian, 1: the 2nd-order star"""
    b = _f(b)
    nx, ny, nz = b.shape
    x = np.zeros_like(b, order="F")
    hist = np.zeros(maxit + 1)
    rnorm = ctypes.c_double(0)
    reason = ctypes.c_int(0)
    its = LIB.orc_cg_solve_op(op, nx, ny, nz, _dx(dx), _p(b), _p(x), rtol, abstol, maxit,
                           ctypes.byref(rnorm), ctypes.byref(reason), _p(hist), len(hist))
    return x, its, rnorm.value, reason.value, hist[: its + 1]
