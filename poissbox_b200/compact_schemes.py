"""Host-side mirror of the reference's `module compact_schemes` (src/compact_schemes.f90:9-13):
same names, argument meaning and error behaviour, numpy arrays in Fortran order f(i,j,k) in and
out.  Every function is one call of a `pbx_*_host` entry point of the C ABI (include/pbx.h), which
stages the data to the GPU and back -- this is the path the reference's own tests exercise after
the drop-in.  A length mismatch raises SizeMismatch (code 7, the reference's
`stop 7`)."""
import numpy as np

from . import _lib
from ._lib import LIB, MODE_FAST, MODE_REFERENCE, check

_dp = _lib._dp


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def _poisoned(shape):
    # the reference's tests pre-fill outputs with 73.29 (tests/lapl/test_lapl.f90:62)
    return np.full(shape, 73.29, order="F")


def set_host_mode(mode):
    """schedule of grad / div / interp below (they have no mode argument in the reference):
    MODE_REFERENCE (default, bit-identical to the reference's order of operations) or MODE_FAST"""
    check(LIB.pbx_host_set_mode(int(mode)))


def lapl(f, dx, mode=MODE_FAST):
    """compact_schemes::lapl, src/compact_schemes.f90:17-37"""
    f = _f(f)
    nx, ny, nz = f.shape
    out = _poisoned(f.shape)
    check(LIB.pbx_lapl_host(nx, ny, nz, _p(f), _lib._d3(*[float(v) for v in dx]), _p(out), mode))
    return out


def lapl_batch(fields, dx, mode=MODE_FAST):
    """compact_schemes::lapl of several fields of one box in one call (pbx_lapl_host_batch): copies in,
    compute and copies out of consecutive fields overlap on three streams"""
    import ctypes

    fs = [_f(f) for f in fields]
    nx, ny, nz = fs[0].shape
    if any(f.shape != (nx, ny, nz) for f in fs):
        raise ValueError("lapl_batch expects fields of one shape")
    outs = [_poisoned((nx, ny, nz)) for _ in fs]
    n = len(fs)
    pin = (ctypes.c_void_p * n)(*[f.ctypes.data for f in fs])
    pout = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
    check(LIB.pbx_lapl_host_batch(nx, ny, nz, n, pin, _lib._d3(*[float(v) for v in dx]), pout, mode))
    return outs


def grad(f, dx):
    """compact_schemes::grad, :42-88 -> df(nx,ny,nz,3)"""
    f = _f(f)
    nx, ny, nz = f.shape
    out = _poisoned((nx, ny, nz, 3))
    check(LIB.pbx_grad_host(nx, ny, nz, _p(f), _lib._d3(*[float(v) for v in dx]), _p(out)))
    return out


def div(f, dx):
    """compact_schemes::div, :207-257; f(nx,ny,nz,3)"""
    f = _f(f)
    nx, ny, nz, nc = f.shape
    if nc != 3:
        raise ValueError("div expects a 3-component field")
    out = _poisoned((nx, ny, nz))
    check(LIB.pbx_div_host(nx, ny, nz, _p(f), _lib._d3(*[float(v) for v in dx]), _p(out)))
    return out


def interp(f, opt_stagger=-1):
    """compact_schemes::interp, :93-142"""
    f = _f(f)
    nx, ny, nz = f.shape
    out = _poisoned(f.shape)
    check(LIB.pbx_interp_host(nx, ny, nz, _p(f), _p(out), int(opt_stagger)))
    return out


def interp_div(f):
    """compact_schemes::interp_div, :144-152"""
    return interp(f, +1)


def grad_1d(f, dx, df=None, opt_stagger=-1):
    """compact_schemes::grad_1d, :155-204.  `df` (optional) is the caller's output array, whose
    length is checked against f as the reference does."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    df = _poisoned(f.shape) if df is None else df
    check(LIB.pbx_grad_1d_host(len(f), _p(f), float(dx), len(df), _p(df), int(opt_stagger)))
    return df


def div_1d(f, dx, df=None):
    """compact_schemes::div_1d, :260-268"""
    return grad_1d(f, dx, df, +1)


def interp_1d(f, fi=None, opt_stagger=-1):
    """compact_schemes::interp_1d, :271-319"""
    f = np.ascontiguousarray(f, dtype=np.float64)
    fi = _poisoned(f.shape) if fi is None else fi
    check(LIB.pbx_interp_1d_host(len(f), _p(f), len(fi), _p(fi), int(opt_stagger)))
    return fi


def interp_1d_div(f, fi=None):
    """compact_schemes::interp_1d_div, :322-329"""
    return interp_1d(f, fi, +1)
