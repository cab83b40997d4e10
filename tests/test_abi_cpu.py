"""CPU: the C-ABI library loads, exports every symbol include/pbx.h declares, validates arguments,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import poissbox_b200 as pbx
from poissbox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pbx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pbx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    syms = header_symbols()
    assert len(syms) >= 39
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pbx.h but not exported by libpbx.so"


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_strings():
    assert pbx.LIB.pbx_version() == 100
    assert pbx.LIB.pbx_error_string(7) == b"array size mismatch"
    assert pbx.LIB.pbx_error_string(0) == b"success"


def test_library_is_self_contained():
    """no torch, no oracle, no NCCL link-time dependency in the product library"""
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "nccl" not in out


def test_size_mismatch_is_stop_7():
    """src/compact_schemes.f90:177-180: the length check comes before anything else"""
    f = np.zeros(8)
    with pytest.raises(pbx.SizeMismatch) as e:
        pbx.compact_schemes.grad_1d(f, 0.1, df=np.zeros(7))
    assert e.value.code == 7
    with pytest.raises(pbx.SizeMismatch):
        pbx.compact_schemes.interp_1d(f, fi=np.zeros(9))


def test_bad_arguments():
    h = ctypes.c_void_p()
    dx = _lib._d3(1.0, 1.0, 1.0)
    assert pbx.LIB.pbx_create(2, 16, 16, dx, 0, None, ctypes.byref(h)) == _lib.PBX_ERR_ARG
    assert pbx.LIB.pbx_create(16, 16, 16, _lib._d3(1.0, -1.0, 1.0), 0, None, ctypes.byref(h)) == _lib.PBX_ERR_ARG
    assert pbx.LIB.pbx_lapl_device(None, None, None) == _lib.PBX_ERR_ARG
    assert pbx.LIB.pbx_set_mode(None, 0) == _lib.PBX_ERR_ARG
    assert pbx.LIB.pbx_destroy(None) == 0


def _no_gpu():
    return pbx.LIB.pbx_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    f = np.zeros((16, 16, 16))
    with pytest.raises(pbx.PbxError) as e:
        pbx.compact_schemes.lapl(f, [1, 1, 1])
    assert e.value.code == _lib.PBX_ERR_CUDA
    with pytest.raises(pbx.PbxError):
        pbx.tridsol.tdma(np.ones(4), np.ones(4) * 4, np.ones(4), np.ones(4))
    with pytest.raises(pbx.PbxError):
        pbx.Handle(16, 16, 16, (1, 1, 1))


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under poissbox_b200/, bench.py's own arm excluded,
    may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poissbox_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle_lib" not in txt and "pbx_oracle" not in txt and "orc_" not in txt, fn
