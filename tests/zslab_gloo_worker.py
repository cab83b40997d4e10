"""worker of test_zslab_cpu.py::test_slab_protocol_gloo_world2 (one process per rank, gloo)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import zslab_model as zm

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
nzl, dz, nl = 64, 1.0 / 128, 7
rng = np.random.default_rng(42)                      # same stream on every rank: the global fields
c = rng.uniform(-1, 1, (world * nzl, nl))
d = rng.uniform(-1, 1, (world * nzl, nl))
cl, dl = c[rank * nzl:(rank + 1) * nzl], d[rank * nzl:(rank + 1) * nzl]
up, dn = (np.stack(m) for m in zm.boundary_messages(cl, dl, dz))     # [9, nl] each
lower, upper = (rank - 1) % world, (rank + 1) % world
recv_lo, recv_up = torch.zeros(up.shape, dtype=torch.float64), torch.zeros(dn.shape, dtype=torch.float64)
# the exchange of the CUDA path (pbx_dist.cu: dist_exchange_nccl): send up / send down, receive from below / above
ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(up)), upper, tag=0),
       dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(dn)), lower, tag=1),
       dist.P2POp(dist.irecv, recv_lo, lower, tag=0),
       dist.P2POp(dist.irecv, recv_up, upper, tag=1)]
for r in dist.batch_isend_irecv(ops):
    r.wait()
out = zm.slab_zpass(cl, dl, dz, list(recv_lo.numpy()), list(recv_up.numpy()))
truth = zm.periodic_truth(c, d, dz)[rank * nzl:(rank + 1) * nzl]
err = np.max(np.abs(out - truth)) / np.max(np.abs(truth))
# a global dot product the way the CG does it: local partial, then all-reduce
loc = torch.tensor([float(np.sum(out * cl))], dtype=torch.float64)
dist.all_reduce(loc)
ref = float(np.sum(zm.periodic_truth(c, d, dz) * c))
ok = err <= 5e-15 and abs(loc.item() - ref) <= 1e-12 * abs(ref)
dist.barrier()
dist.destroy_process_group()
print("ZSLAB_OK" if ok else f"ZSLAB_FAIL err={err} dot={loc.item()} ref={ref}")
sys.exit(0 if ok else 1)
