"""Host-side mirror of the reference's `module tridsol` (src/tridsol.f90:16-18).  Dummy names as
in the reference: a = sub-diagonal, b = DIAGONAL, c = super-diagonal, d = rhs/solution.  Arrays
declared intent(inout) there are updated in place here too (and returned)."""
import numpy as np

from . import _lib
from ._lib import LIB, check

_dp = _lib._dp


def _chk(*arrs):
    for a in arrs:
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 1):
            raise ValueError("tridsol routines take contiguous 1-D float64 numpy arrays")
    if len({len(a) for a in arrs}) != 1:
        raise ValueError("tridsol arrays must have equal length")
    return len(arrs[0])


def _p(a):
    return a.ctypes.data_as(_dp)


def tdma(a, b, c, d):
    """tridsol::tdma, src/tridsol.f90:22-32 (b is overwritten with the pivots, d with the solution)"""
    n = _chk(a, b, c, d)
    check(LIB.pbx_tdma_host(n, _p(a), _p(b), _p(c), _p(d)))
    return d


def tdma_periodic(a, b, c, d):
    """tridsol::tdma_periodic, :34-74 (b untouched, d overwritten)"""
    n = _chk(a, b, c, d)
    check(LIB.pbx_tdma_periodic_host(n, _p(a), _p(b), _p(c), _p(d)))
    return d


def fwd_sweep(a, b, c, d):
    """tridsol::fwd_sweep, :76-96"""
    n = _chk(a, b, c, d)
    check(LIB.pbx_fwd_sweep_host(n, _p(a), _p(b), _p(c), _p(d)))
    return b, d


def bwd_sweep(b, c, d):
    """tridsol::bwd_sweep, :98-115"""
    n = _chk(b, c, d)
    check(LIB.pbx_bwd_sweep_host(n, _p(b), _p(c), _p(d)))
    return d
