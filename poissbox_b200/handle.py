"""Device-resident operator handle: the counterpart of the reference's shell-matrix context
`mat_ctx` (src/poissbox.f90:17-20), its MatMult callback `mfmult` (:300-322) and `solve` (:269-298).

Fields are torch float64 CUDA tensors whose memory is Fortran column-major f(i,j,k), i.e. a
C-contiguous tensor of shape (nz, ny, nx) (or (3, nz, ny, nx) for vector fields).  torch is used
for device memory and streams only; every operation is a call into libpbx.so.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import LIB, check


def fortran_to_torch(a, device="cuda"):
    """numpy array f(i,j,k[,c]) in any order -> torch tensor (..., nz, ny, nx) with the same memory
    layout as the Fortran array."""
    import torch

    a = np.asarray(a, dtype=np.float64)
    return torch.from_numpy(np.ascontiguousarray(a.transpose(*reversed(range(a.ndim))))).to(device)


def torch_to_fortran(t):
    """inverse of fortran_to_torch: returns a Fortran-ordered numpy array f(i,j,k[,c])"""
    a = t.detach().cpu().numpy()
    return np.asfortranarray(a.transpose(*reversed(range(a.ndim))))


class Handle:
    def __init__(self, nx, ny, nz, dx, device=0, comm=None, slab=None):
        """nz is the LOCAL number of planes.  comm: an ncclComm_t (int) for the z-slab
        decomposition over NCCL; slab=(rank, nranks): a phase-driven slab handle (no NCCL)."""
        import torch

        self._torch = torch
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.dx = tuple(float(v) for v in dx)
        self.device = int(device)
        self._h = ctypes.c_void_p()
        if slab is not None:
            check(LIB.pbx_create_slab(self.nx, self.ny, self.nz, _lib._d3(*self.dx), self.device,
                                      int(slab[0]), int(slab[1]), ctypes.byref(self._h)))
        else:
            check(LIB.pbx_create(self.nx, self.ny, self.nz, _lib._d3(*self.dx), self.device,
                                 ctypes.c_void_p(comm) if comm else None, ctypes.byref(self._h)))

    # -- phase-driven z-slab decomposition (single-process emulation; see include/pbx.h) ----------
    def slab_phase1(self, f):
        check(LIB.pbx_slab_phase1(self._h, self._field(f)))

    def slab_phase2(self, out=None):
        out = self.empty() if out is None else out
        check(LIB.pbx_slab_phase2(self._h, self._field(out)))
        return out

    def slab_op_phase1(self, op, f):
        """grad / div / interp on a slab (op = _lib.OP_*): stages before the z operators + boundary sweeps"""
        check(LIB.pbx_slab_op_phase1(self._h, int(op), self._field(f, 3 if op == _lib.OP_DIV else 1)))

    def slab_op_phase2(self, op, f, out=None):
        if out is None:
            out = self.empty(3 if op == _lib.OP_GRAD else 1)
        check(LIB.pbx_slab_op_phase2(self._h, int(op), self._field(f, 3 if op == _lib.OP_DIV else 1),
                                     self._field(out, 3 if op == _lib.OP_GRAD else 1)))
        return out

    def slab_get_messages(self):
        """the two outgoing messages of phase 1 (to rank+1, to rank-1) for a host-owned exchange"""
        n = ctypes.c_longlong()
        check(LIB.pbx_slab_message_count(self._h, ctypes.byref(n)))
        up = self._torch.empty(n.value, dtype=self._torch.float64, device=f"cuda:{self.device}")
        dn = self._torch.empty_like(up)
        check(LIB.pbx_slab_get_messages(self._h, ctypes.c_void_p(up.data_ptr()), ctypes.c_void_p(dn.data_ptr())))
        return up, dn

    def slab_put_messages(self, from_lo, from_up):
        check(LIB.pbx_slab_put_messages(self._h, ctypes.c_void_p(from_lo.data_ptr()),
                                        ctypes.c_void_p(from_up.data_ptr())))

    @staticmethod
    def slab_exchange_local(handles):
        arr = (ctypes.c_void_p * len(handles))(*[h._h for h in handles])
        check(LIB.pbx_slab_exchange_local(arr, len(handles)))

    @staticmethod
    def slab_link_local(handles):
        """Peer boards (include/pbx.h, pbx_slab_link_peers) for a ring of slab handles living in ONE
        process on one device (or peer-enabled devices): afterwards every handle works like one
        created with a communicator -- lapl / grad / div / interp / cg_solve exchange and reduce
        among themselves on the device.  Each handle needs its own stream (set_stream) and its own
        host thread, because its kernels wait on the device for the other ranks'."""
        n = len(handles)
        bufs = []
        for h in handles:
            p = ctypes.c_void_p()
            check(LIB.pbx_slab_recv_buffer(h._h, ctypes.byref(p)))
            bufs.append(p.value)
        arr = (ctypes.c_void_p * n)(*bufs)
        for h in handles:
            check(LIB.pbx_slab_link_peers(h._h, arr, n))

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self):
        if self._h:
            LIB.pbx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def mode(self):
        m = ctypes.c_int()
        check(LIB.pbx_get_mode(self._h, ctypes.byref(m)))
        return m.value

    @mode.setter
    def mode(self, m):
        check(LIB.pbx_set_mode(self._h, int(m)))

    def use_current_stream(self):
        check(LIB.pbx_set_stream(self._h, ctypes.c_void_p(self._torch.cuda.current_stream().cuda_stream)))

    def set_stream(self, stream_ptr):
        check(LIB.pbx_set_stream(self._h, ctypes.c_void_p(stream_ptr)))

    def synchronize(self):
        check(LIB.pbx_synchronize(self._h))

    @property
    def launches(self):
        return int(LIB.pbx_launch_count(self._h))

    # -- helpers --------------------------------------------------------------------------------
    def _field(self, t, ncomp=1):
        torch = self._torch
        n = self.nx * self.ny * self.nz * ncomp
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == n):
            raise ValueError(f"expected a contiguous float64 CUDA tensor with {n} elements")
        return ctypes.c_void_p(t.data_ptr())

    def empty(self, ncomp=1):
        shape = (self.nz, self.ny, self.nx) if ncomp == 1 else (ncomp, self.nz, self.ny, self.nx)
        return self._torch.empty(shape, dtype=self._torch.float64, device=f"cuda:{self.device}")

    # -- operators (compact_schemes) ------------------------------------------------------------
    def lapl(self, f, out=None):
        out = self.empty() if out is None else out
        check(LIB.pbx_lapl_device(self._h, self._field(f), self._field(out)))
        return out

    def mult(self, x, y=None):
        """MatMult of the shell matrix: y = A x (mfmult, src/poissbox.f90:300-322) for the handle's
        operator (`operator` property: the compact Laplacian, or the 2nd-order star mfmult applies today)"""
        y = self.empty() if y is None else y
        check(LIB.pbx_matmult_device(self._h, self._field(x), self._field(y)))
        return y

    def star(self, x, out=None):
        """the 2nd-order 7-point star (compute_lapl_pointwise, src/poissbox.f90:84-126)"""
        out = self.empty() if out is None else out
        check(LIB.pbx_star_device(self._h, self._field(x), self._field(out)))
        return out

    @property
    def operator(self):
        op = ctypes.c_int()
        check(LIB.pbx_get_operator(self._h, ctypes.byref(op)))
        return op.value

    @operator.setter
    def operator(self, op):
        check(LIB.pbx_set_operator(self._h, int(op)))

    def lapl_dot(self, f, out=None):
        out = self.empty() if out is None else out
        dot = self._torch.zeros(1, dtype=self._torch.float64, device=f"cuda:{self.device}")
        check(LIB.pbx_lapl_dot_device(self._h, self._field(f), self._field(out), ctypes.c_void_p(dot.data_ptr())))
        return out, dot

    def lapl_profile(self, f, out, reps=5):
        """average ms of the x, y, z pass kernels of the FAST Laplacian (CUDA events)"""
        ms = _lib._d3()
        check(LIB.pbx_lapl_profile_device(self._h, self._field(f), self._field(out), int(reps), ms))
        return tuple(ms)

    def grad(self, f, out=None):
        out = self.empty(3) if out is None else out
        check(LIB.pbx_grad_device(self._h, self._field(f), self._field(out, 3)))
        return out

    def div(self, f, out=None):
        out = self.empty() if out is None else out
        check(LIB.pbx_div_device(self._h, self._field(f, 3), self._field(out)))
        return out

    def interp(self, f, stagger=-1, out=None):
        out = self.empty() if out is None else out
        check(LIB.pbx_interp_device(self._h, self._field(f), self._field(out), int(stagger)))
        return out

    def set_pc(self, pc, nu=0):
        """preconditioner of cg_solve: _lib.PC_NONE (default) or _lib.PC_MG, a V(nu, nu) multigrid
        cycle on the 2nd-order star (the role of `-pc_type gamg` on P, src/poissbox.f90:294)"""
        check(LIB.pbx_set_pc(self._h, int(pc), int(nu)))

    def pc_apply(self, r, z=None):
        z = self.empty() if z is None else z
        check(LIB.pbx_pc_apply_device(self._h, self._field(r), self._field(z)))
        return z

    # -- solve ----------------------------------------------------------------------------------
    def ksp_solve(self, b, options="", x=None):
        """The reference's solve() (src/poissbox.f90:269-298) configured the way the reference configures
        it, by PETSc option names (KSPSetFromOptions :295; README.md:43-49), e.g.
        "-ksp_type cg -pc_type gamg -ksp_rtol 1e-8 -ksp_monitor -ksp_converged_reason".
        Returns (x, its, rnorm, reason)."""
        x = self.empty() if x is None else x
        its, reason, rnorm = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        check(LIB.pbx_ksp_solve_device(self._h, options.encode(), self._field(b), self._field(x),
                                       ctypes.byref(its), ctypes.byref(rnorm), ctypes.byref(reason)))
        return x, its.value, rnorm.value, reason.value

    def cg_solve(self, b, x=None, rtol=1e-5, abstol=1e-50, maxit=10000):
        """KSPSolve with -ksp_type cg -pc_type none (src/poissbox.f90:293-296).
        Returns (x, its, rnorm, reason, history)."""
        x = self.empty() if x is None else x
        its, reason, rnorm = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        hist = np.zeros(maxit + 1)
        check(LIB.pbx_cg_solve_device(self._h, self._field(b), self._field(x), rtol, abstol, int(maxit),
                                      ctypes.byref(its), ctypes.byref(rnorm), ctypes.byref(reason),
                                      hist.ctypes.data_as(_lib._dp), len(hist)))
        return x, its.value, rnorm.value, reason.value, hist[: its.value + 1]
