import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# The operator parity files run first, the solver tests after them, so that nothing downstream of the
# operators (a CG heuristic, a preconditioner) can keep `pytest -x` from reaching the parity evidence.
_ORDER = ["test_parity_gpu", "test_golden", "test_zslab_gpu", "test_cpp_mirror", "test_petsc_glue", "test_cg_gpu"]


def _rank(item):
    name = os.path.basename(str(item.fspath))
    for k, stem in enumerate(_ORDER):
        if name.startswith(stem):
            return k
    return len(_ORDER)


def pytest_collection_modifyitems(config, items):
    items.sort(key=_rank)   # stable: the order inside a file is kept
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
