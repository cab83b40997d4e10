"""Worker of tests/test_emu_schedules.py: runs a set of kernels on the CPU kernel-logic harness
under the thread order selected by PBX_EMU_SCHED and prints one digest of all results."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import emu_lib
from poissbox_b200 import _lib


def digest(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    rng = np.random.default_rng(1)
    res = []
    for shape in [(64, 32, 64), (16, 1024, 16), (1024, 16, 16)]:
        dx = tuple(1.0 / n for n in shape)
        f = np.asfortranarray(rng.uniform(-1, 1, shape))
        for no_tma in ("0", "1"):
            os.environ["PBX_NO_TMA"] = no_tma
            h = emu_lib.EmuHandle(*shape, dx)
            out, dot = h.lapl_dot(f)
            res += [digest(out), repr(dot)]
            if no_tma == "0" and shape[0] == 64:
                v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
                res += [digest(h.grad(f)), digest(h.div(v)), digest(h.interp(f))]
                h.set_mode(1)
                res.append(digest(h.lapl(f)))
            h.close()
    os.environ.pop("PBX_NO_TMA")
    shape, P, nzl, dx = (32, 16, 128), 2, 64, (1 / 32, 0.7 / 16, 1.3 / 128)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    slabs = [emu_lib.EmuHandle(32, 16, nzl, dx, slab=(r, P)) for r in range(P)]
    for r, h in enumerate(slabs):
        h.slab_phase1(np.asfortranarray(f[:, :, r * nzl:(r + 1) * nzl]))
    emu_lib.EmuHandle.slab_exchange_local(slabs)
    res.append(digest(np.concatenate([h.slab_phase2() for h in slabs], axis=2)))
    for op in (_lib.OP_GRAD, _lib.OP_DIV):
        src = np.asfortranarray(rng.uniform(-1, 1, shape + (3,))) if op == _lib.OP_DIV else f
        for r, h in enumerate(slabs):
            h.slab_op_phase1(op, np.asfortranarray(src[:, :, r * nzl:(r + 1) * nzl]))
        emu_lib.EmuHandle.slab_exchange_local(slabs)
        res.append(digest(np.concatenate([h.slab_op_phase2(op) for h in slabs], axis=2)))
    n = 16
    dx = (2 * np.pi / n,) * 3
    b = np.asfortranarray(rng.uniform(-1, 1, (n, n, n)))
    b -= b.mean()
    h = emu_lib.EmuHandle(n, n, n, dx)
    x, its, _, _, _ = h.cg_solve(b, rtol=1e-6)
    res.append(f"{its}:{digest(x)}")
    h.set_pc(_lib.PC_MG, 2)
    x, its, _, _, _ = h.cg_solve(b, rtol=1e-6)
    res.append(f"{its}:{digest(x)}")
    # grad / div through the TMA-pipelined line operators, and the swizzled y/z tiles of the Laplacian
    os.environ["PBX_LINEOP_TMA"] = "1"
    os.environ["PBX_YZ_ROT"] = "1"
    shape = (64, 32, 64)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
    h = emu_lib.EmuHandle(*shape, tuple(1.0 / m for m in shape))
    res += [digest(h.grad(f)), digest(h.div(v)), digest(h.lapl(f))]
    h.close()
    os.environ.pop("PBX_LINEOP_TMA")
    os.environ.pop("PBX_YZ_ROT")
    h = emu_lib.EmuHandle(16, 16, 16, (0.1,) * 3)
    # line-major tridiagonal batches through the TMA tiles (pbx_tdma_tma.cu)
    os.environ["PBX_TDMA_TMA"] = "1"
    n, nl = 40, 45
    for per in (0, 1):
        a, c = rng.uniform(-0.3, 0.3, (nl, n)), rng.uniform(-0.3, 0.3, (nl, n))
        bd, d = 1.0 + rng.uniform(0, 0.5, (nl, n)), rng.uniform(-1, 1, (nl, n))
        arrs = [emu_lib.aligned(np.asfortranarray(v.T)) for v in (a, bd, c, d)]    # memory: [line][i]
        fn = h.lib.pbx_tdma_periodic_batch_device if per else h.lib.pbx_tdma_batch_device
        emu_lib.check(h.lib, fn(n, nl, 1, n, *[emu_lib.ptr(v) for v in arrs], None))
        res += [digest(arrs[1]), digest(arrs[3])]
    os.environ.pop("PBX_TDMA_TMA")
    print(hashlib.md5(" ".join(res).encode()).hexdigest())


if __name__ == "__main__":
    main()
