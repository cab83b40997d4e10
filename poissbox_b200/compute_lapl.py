"""Host-side mirror of the reference's `module compute_lapl` (src/poissbox.f90:75-150): the
2nd-order 7-point star its MATSHELL callback `mfmult` applies today, on a periodic box (the DMDA's
periodic ghosts), numpy arrays in Fortran order in and out.  One call of `pbx_star_host`."""
import numpy as np

from . import _lib
from ._lib import LIB, check

_dp = _lib._dp


def compute_lapl_pointwise(x, grid_deltas):
    """compute_lapl::compute_lapl_pointwise, src/poissbox.f90:84-126: b_i = stencil_op(x, i)"""
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    nx, ny, nz = x.shape
    b = np.full(x.shape, 73.29, order="F")
    check(LIB.pbx_star_host(nx, ny, nz, x.ctypes.data_as(_dp), _lib._d3(*[float(v) for v in grid_deltas]),
                            b.ctypes.data_as(_dp)))
    return b
