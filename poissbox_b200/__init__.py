"""poissbox-b200: the B200 (sm_100a) implementation of the compact-scheme Laplacian hot path of
3decomp/poissbox.  The product is the C-ABI library (include/pbx.h, poissbox_b200/csrc); this
package is the host-side mirror of the reference's module interfaces on top of it:

  poissbox_b200.compact_schemes   <->  module compact_schemes  (src/compact_schemes.f90:9-13)
  poissbox_b200.tridsol           <->  module tridsol          (src/tridsol.f90:16-18)
  poissbox_b200.compute_lapl      <->  module compute_lapl     (src/poissbox.f90:75-150)
  poissbox_b200.Handle            <->  mat_ctx + mfmult + solve (src/poissbox.f90:17-20,269-322)
"""
from ._lib import LIB, LIB_PATH, MODE_FAST, MODE_REFERENCE, PbxError, SizeMismatch, check
from .handle import Handle, fortran_to_torch, torch_to_fortran
from . import compact_schemes, compute_lapl, tridsol

__all__ = ["LIB", "LIB_PATH", "MODE_FAST", "MODE_REFERENCE", "PbxError", "SizeMismatch", "check",
           "Handle", "fortran_to_torch", "torch_to_fortran", "compact_schemes", "compute_lapl", "tridsol"]
