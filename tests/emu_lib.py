"""Loader of the CPU kernel-logic harness (tests/emu): the library's own kernel sources run on a
fiber model of the CUDA execution model, behind the same C ABI.  TEST INFRASTRUCTURE ONLY -- used
by the `-m "not gpu"` tests to check kernel logic (indexing, tiles, barriers, look-back, slab
boundaries) where there is no GPU; host numpy arrays stand in for device memory.  The product
never loads it."""
import ctypes
import os
import subprocess

import numpy as np

from poissbox_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(_HERE, "emu")
EMU_PATH = os.path.join(EMU_DIR, "_build", "libpbx_emu.so")

_cached = None


def load():
    global _cached
    if _cached is None:
        subprocess.check_call(["make", "-s", "-j8", "-C", EMU_DIR], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(EMU_PATH)
        for name, (res, args) in _lib.SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        lib.pbx_emu_launches_total.restype = ctypes.c_longlong
        lib.pbx_emu_tensor_maps_total.restype = ctypes.c_longlong
        lib.pbx_emu_tensor_maps3_swizzled_total.restype = ctypes.c_longlong
        _cached = lib
    return _cached


def check(lib, rc):
    if rc != 0:
        raise _lib.PbxError(rc, lib.pbx_last_error().decode() or lib.pbx_error_string(rc).decode())


def ptr(a):
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return ctypes.c_void_p(a.ctypes.data)


def new_field(shape):
    """16-byte aligned Fortran-ordered array pre-filled with the reference tests' poison value"""
    n = int(np.prod(shape))
    raw = np.empty(n + 2)
    off = (-(raw.ctypes.data // 8)) % 2
    a = raw[off:off + n].reshape(shape, order="F")
    a[...] = 73.29
    return a


def aligned(a):
    out = new_field(a.shape)
    out[...] = a
    return out


class EmuHandle:
    """pbx_create / pbx_*_device on the harness, numpy arrays as device memory"""

    def __init__(self, nx, ny, nz, dx, slab=None):
        self.lib = load()
        self.shape = (nx, ny, nz)
        self._h = ctypes.c_void_p()
        d3 = _lib._d3(*[float(v) for v in dx])
        if slab is None:
            check(self.lib, self.lib.pbx_create(nx, ny, nz, d3, 0, None, ctypes.byref(self._h)))
        else:
            check(self.lib, self.lib.pbx_create_slab(nx, ny, nz, d3, 0, slab[0], slab[1], ctypes.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.pbx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        self.close()

    def set_mode(self, m):
        check(self.lib, self.lib.pbx_set_mode(self._h, m))

    def lapl(self, f):
        f = aligned(f)
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_lapl_device(self._h, ptr(f), ptr(out)))
        return out

    def star(self, f):
        f = aligned(f)
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_star_device(self._h, ptr(f), ptr(out)))
        return out

    def set_operator(self, op):
        check(self.lib, self.lib.pbx_set_operator(self._h, op))

    def mult(self, f):
        f = aligned(f)
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_matmult_device(self._h, ptr(f), ptr(out)))
        return out

    def set_pc(self, pc, nu=0):
        check(self.lib, self.lib.pbx_set_pc(self._h, pc, nu))

    def pc_apply(self, r):
        r = aligned(r)
        z = new_field(self.shape)
        check(self.lib, self.lib.pbx_pc_apply_device(self._h, ptr(r), ptr(z)))
        return z

    def lapl_dot(self, f):
        f = aligned(f)
        out = new_field(self.shape)
        dot = np.zeros(2)
        check(self.lib, self.lib.pbx_lapl_dot_device(self._h, ptr(f), ptr(out), ptr(dot)))
        return out, dot[0]

    def grad(self, f):
        f = aligned(f)
        out = new_field(self.shape + (3,))
        check(self.lib, self.lib.pbx_grad_device(self._h, ptr(f), ptr(out)))
        return out

    def div(self, f):
        f = aligned(f)
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_div_device(self._h, ptr(f), ptr(out)))
        return out

    def interp(self, f, stagger=-1):
        f = aligned(f)
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_interp_device(self._h, ptr(f), ptr(out), stagger))
        return out

    def cg_solve(self, b, rtol=1e-5, abstol=1e-50, maxit=10000):
        b = aligned(b)
        x = new_field(self.shape)
        its, reason, rnorm = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        hist = np.zeros(maxit + 1)
        check(self.lib, self.lib.pbx_cg_solve_device(self._h, ptr(b), ptr(x), rtol, abstol, maxit,
                                                     ctypes.byref(its), ctypes.byref(rnorm),
                                                     ctypes.byref(reason), hist.ctypes.data_as(_lib._dp),
                                                     len(hist)))
        return x, its.value, rnorm.value, reason.value, hist[: its.value + 1]

    def ksp_solve(self, b, options=""):
        b = aligned(b)
        x = new_field(self.shape)
        its, reason, rnorm = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        check(self.lib, self.lib.pbx_ksp_solve_device(self._h, options.encode(), ptr(b), ptr(x), ctypes.byref(its),
                                                      ctypes.byref(rnorm), ctypes.byref(reason)))
        return x, its.value, rnorm.value, reason.value

    def slab_phase1(self, f):
        self._f = aligned(f)
        check(self.lib, self.lib.pbx_slab_phase1(self._h, ptr(self._f)))

    def slab_phase2(self):
        out = new_field(self.shape)
        check(self.lib, self.lib.pbx_slab_phase2(self._h, ptr(out)))
        return out

    def slab_op_phase1(self, op, f):
        self._f = aligned(f)
        check(self.lib, self.lib.pbx_slab_op_phase1(self._h, op, ptr(self._f)))

    def slab_op_phase2(self, op):
        out = new_field(self.shape + (3,) if op == _lib.OP_GRAD else self.shape)
        check(self.lib, self.lib.pbx_slab_op_phase2(self._h, op, ptr(self._f), ptr(out)))
        return out

    def slab_get_messages(self):
        n = ctypes.c_longlong()
        check(self.lib, self.lib.pbx_slab_message_count(self._h, ctypes.byref(n)))
        up, dn = np.empty(n.value), np.empty(n.value)
        check(self.lib, self.lib.pbx_slab_get_messages(self._h, ptr(up), ptr(dn)))
        return up, dn

    def slab_put_messages(self, from_lo, from_up):
        check(self.lib, self.lib.pbx_slab_put_messages(self._h, ptr(from_lo), ptr(from_up)))

    @staticmethod
    def slab_exchange_local(handles):
        lib = handles[0].lib
        arr = (ctypes.c_void_p * len(handles))(*[h._h for h in handles])
        check(lib, lib.pbx_slab_exchange_local(arr, len(handles)))
