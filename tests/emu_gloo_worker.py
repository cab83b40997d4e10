"""worker of test_zslab_cpu.py::test_slab_kernels_gloo_world2: one process per rank (gloo); every
rank runs the library's own slab kernels on the CPU kernel-logic harness (tests/emu), owns the
exchange (pbx_slab_get_messages / torch.distributed send-recv / pbx_slab_put_messages) and checks
its slab of the Laplacian, grad, div and the star against the whole brick computed by rank 0's
means (every rank evaluates the whole brick itself: the global field comes from a shared seed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import emu_lib
from poissbox_b200 import _lib

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
nx, ny, nzl = 32, 16, 64
nz = nzl * world
dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
rng = np.random.default_rng(42)
f = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
v = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz, 3)))
whole = emu_lib.EmuHandle(nx, ny, nz, dx)
mine = slice(rank * nzl, (rank + 1) * nzl)
h = emu_lib.EmuHandle(nx, ny, nzl, dx, slab=(rank, world))
lower, upper = (rank - 1) % world, (rank + 1) % world


def exchange():
    up, dn = h.slab_get_messages()
    recv_lo, recv_up = torch.zeros(len(up), dtype=torch.float64), torch.zeros(len(dn), dtype=torch.float64)
    ops = [dist.P2POp(dist.isend, torch.from_numpy(up), upper, tag=0),
           dist.P2POp(dist.isend, torch.from_numpy(dn), lower, tag=1),
           dist.P2POp(dist.irecv, recv_lo, lower, tag=0),
           dist.P2POp(dist.irecv, recv_up, upper, tag=1)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    h.slab_put_messages(recv_lo.numpy(), recv_up.numpy())


errs = {}
h.slab_phase1(np.asfortranarray(f[:, :, mine]))
exchange()
out = h.slab_phase2()
ref = whole.lapl(f)
errs["lapl"] = np.max(np.abs(out - ref[:, :, mine])) / np.max(np.abs(ref))
for name, op, src, want in (("grad", _lib.OP_GRAD, f, whole.grad(f)), ("div", _lib.OP_DIV, v, whole.div(v)),
                            ("interp", _lib.OP_INTERP, f, whole.interp(f)), ("star", _lib.OP_STAR, f, whole.star(f))):
    h.slab_op_phase1(op, np.asfortranarray(src[:, :, mine]))
    exchange()
    got = h.slab_op_phase2(op)
    errs[name] = np.max(np.abs(got - want[:, :, mine])) / np.max(np.abs(want))
# a global dot product the way the CG does it: local partial, then all-reduce
loc = torch.tensor([float(np.vdot(f[:, :, mine], out))], dtype=torch.float64)
dist.all_reduce(loc)
ref_dot = float(np.vdot(f, ref))
ok = all(e <= 1e-13 for e in errs.values()) and errs["star"] == 0.0 and abs(loc.item() - ref_dot) <= 1e-12 * abs(ref_dot)
dist.barrier()
dist.destroy_process_group()
print(("EMU_GLOO_OK " if ok else "EMU_GLOO_FAIL ") + str(errs))
sys.exit(0 if ok else 1)
