"""ctypes binding of libpbx.so, the C-ABI library declared in include/pbx.h.

The shared object is built in-tree by ``__graft_entry__.build()`` (or ``make -C
poissbox_b200/csrc``).  There is no fallback of any kind: if the library is missing the import
fails, and every compute call fails with ``PbxError`` when no CUDA device is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpbx.so")

c_int, c_ll, c_double, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p
_dp = ctypes.POINTER(c_double)
_ip = ctypes.POINTER(c_int)
_d3 = c_double * 3

PBX_OK, PBX_ERR_ARG, PBX_ERR_CUDA, PBX_ERR_NCCL, PBX_ERR_UNSUPPORTED, PBX_ERR_NOMEM = 0, 1, 2, 3, 4, 5
PBX_ERR_SIZE = 7
MODE_FAST, MODE_REFERENCE = 0, 1
OP_GRAD, OP_DIV, OP_INTERP, OP_INTERP_DIV, OP_STAR = 1, 2, 3, 4, 5
OPERATOR_COMPACT, OPERATOR_STAR = 0, 1
PC_NONE, PC_MG = 0, 1

# every symbol include/pbx.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pbx_version": (c_int, []),
    "pbx_error_string": (ctypes.c_char_p, [c_int]),
    "pbx_last_error": (ctypes.c_char_p, []),
    "pbx_device_count": (c_int, []),
    "pbx_create": (c_int, [c_int, c_int, c_int, _d3, c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    "pbx_destroy": (c_int, [c_void_p]),
    "pbx_set_mode": (c_int, [c_void_p, c_int]),
    "pbx_get_mode": (c_int, [c_void_p, _ip]),
    "pbx_set_stream": (c_int, [c_void_p, c_void_p]),
    "pbx_synchronize": (c_int, [c_void_p]),
    "pbx_get_dims": (c_int, [c_void_p, _ip, _ip, _ip]),
    "pbx_launch_count": (c_ll, [c_void_p]),
    "pbx_comm_unique_id": (c_int, [c_void_p]),
    "pbx_comm_init_rank": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "pbx_comm_destroy": (c_int, [c_void_p]),
    "pbx_create_slab": (c_int, [c_int, c_int, c_int, _d3, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "pbx_slab_phase1": (c_int, [c_void_p, c_void_p]),
    "pbx_slab_phase2": (c_int, [c_void_p, c_void_p]),
    "pbx_slab_exchange_local": (c_int, [ctypes.POINTER(c_void_p), c_int]),
    "pbx_slab_op_phase1": (c_int, [c_void_p, c_int, c_void_p]),
    "pbx_slab_op_phase2": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "pbx_slab_message_count": (c_int, [c_void_p, ctypes.POINTER(c_ll)]),
    "pbx_slab_get_messages": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_slab_put_messages": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_slab_exchange": (c_int, [c_void_p]),
    "pbx_slab_recv_bytes": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_size_t)]),
    "pbx_slab_recv_buffer": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "pbx_slab_link_peers": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int]),
    "pbx_peer_sync_active": (c_int, [c_void_p]),
    "pbx_allreduce_sum": (c_int, [c_void_p, c_void_p, c_int]),
    "pbx_dist_tables_host": (c_int, [c_int, c_double, _ip, _ip, _ip, _dp, _dp, _dp, _dp, _dp]),
    "pbx_lapl_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_lapl_dot_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbx_lapl_profile_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int, _d3]),
    "pbx_grad_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_div_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_interp_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "pbx_set_operator": (c_int, [c_void_p, c_int]),
    "pbx_get_operator": (c_int, [c_void_p, _ip]),
    "pbx_matmult_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_star_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_star_host": (c_int, [c_int, c_int, c_int, _dp, _d3, _dp]),
    "pbx_grad_1d_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_double, c_void_p, c_int, c_void_p]),
    "pbx_interp_1d_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_int, c_void_p]),
    "pbx_tdma_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbx_tdma_periodic_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbx_fwd_sweep_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbx_bwd_sweep_batch_device": (c_int, [c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbx_cg_solve_device": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, _ip, _dp, _ip, _dp, c_int]),
    "pbx_ksp_solve_device": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, c_void_p, _ip, _dp, _ip]),
    "pbx_set_pc": (c_int, [c_void_p, c_int, c_int]),
    "pbx_pc_apply_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pbx_lapl_host": (c_int, [c_int, c_int, c_int, _dp, _d3, _dp, c_int]),
    "pbx_lapl_host_batch": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p), _d3,
                                    ctypes.POINTER(c_void_p), c_int]),
    "pbx_grad_host": (c_int, [c_int, c_int, c_int, _dp, _d3, _dp]),
    "pbx_div_host": (c_int, [c_int, c_int, c_int, _dp, _d3, _dp]),
    "pbx_interp_host": (c_int, [c_int, c_int, c_int, _dp, _dp, c_int]),
    "pbx_grad_1d_host": (c_int, [c_int, _dp, c_double, c_int, _dp, c_int]),
    "pbx_interp_1d_host": (c_int, [c_int, _dp, c_int, _dp, c_int]),
    "pbx_tdma_host": (c_int, [c_int, _dp, _dp, _dp, _dp]),
    "pbx_tdma_periodic_host": (c_int, [c_int, _dp, _dp, _dp, _dp]),
    "pbx_fwd_sweep_host": (c_int, [c_int, _dp, _dp, _dp, _dp]),
    "pbx_bwd_sweep_host": (c_int, [c_int, _dp, _dp, _dp]),
    "pbx_cg_solve_host": (c_int, [c_int, c_int, c_int, _d3, _dp, _dp, c_double, c_double, c_int, c_int, _ip, _dp, _ip, _dp, c_int]),
    "pbx_host_set_mode": (c_int, [c_int]),
    "pbx_host_cache_clear": (c_int, []),
}


class PbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pbx error {code}: {msg}")
        self.code = code


class SizeMismatch(PbxError):
    """PBX_ERR_SIZE: the reference's `stop 7` (src/compact_schemes.f90:177-180)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C poissbox_b200/csrc` (there is no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


LIB = _load()


def check(rc):
    if rc == PBX_OK:
        return
    detail = LIB.pbx_last_error().decode() or LIB.pbx_error_string(rc).decode()
    if rc == PBX_ERR_SIZE:
        raise SizeMismatch(rc, detail)
    raise PbxError(rc, detail)
