"""torchrun worker: the distributed Laplacian (and its fused dot) on a SEQUENCE of different fields against one handle on
the whole brick -- consecutive MatMults with different inputs are what a CG does, and what shows a slab z pass that
reads a neighbour's message of the previous MatMult (a static input hides it: the stale message is the right one).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        tools/dist_dyn_check.py n NZ [reps]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import poissbox_b200 as pbx

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, NZ = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
nzl = NZ // world
idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    raw = (ctypes.c_ubyte * 128)()
    pbx.check(pbx.LIB.pbx_comm_unique_id(raw))
    idbuf = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
dist.broadcast(idbuf, 0)
raw = (ctypes.c_ubyte * 128)(*idbuf.cpu().tolist())
comm = ctypes.c_void_p()
pbx.check(pbx.LIB.pbx_comm_init_rank(raw, world, rank, local, ctypes.byref(comm)))
dx = (1.0 / n,) * 3
whole = pbx.Handle(n, n, NZ, dx, device=local)
h = pbx.Handle(n, n, nzl, dx, device=local, comm=comm.value)
sl = slice(rank * nzl, (rank + 1) * nzl)
g = torch.Generator(device=dev).manual_seed(99)          # same seed on every rank: the same global fields
fields = [torch.rand((NZ, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1 for _ in range(3)]
refs = [whole.lapl(f) for f in fields]
mine = [f[sl].contiguous() for f in fields]
worst_f, worst_fd, worst_d = 0.0, 0.0, 0.0
# back to back, no synchronisation in between: rep k uses field k % 3
outs = []
for k in range(reps):
    if k % 2 == 0:
        outs.append((k, h.lapl(mine[k % 3]), None))
    else:
        o, d = h.lapl_dot(mine[k % 3])
        outs.append((k, o, d))
torch.cuda.synchronize()
for k, o, d in outs:
    ref = refs[k % 3]
    e = (o - ref[sl]).abs().max().item() / ref.abs().max().item()
    if d is None:
        worst_f = max(worst_f, e)
    else:
        worst_fd = max(worst_fd, e)
        dref = torch.dot(fields[k % 3].flatten(), ref.flatten()).item()
        worst_d = max(worst_d, abs(d.item() - dref) / abs(dref))
ok = worst_f <= 1e-13 and worst_fd <= 1e-13 and worst_d <= 1e-12
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
print(f"rank {rank}/{world}: {reps} MatMults on changing fields: lapl err {worst_f:.2e}, lapl_dot field err {worst_fd:.2e}, "
      f"dot err {worst_d:.2e}  env {sorted((k, v) for k, v in os.environ.items() if k.startswith('PBX_'))} -> {'OK' if ok else 'FAIL'}", flush=True)
h.close()
whole.close()
pbx.LIB.pbx_comm_destroy(comm)
dist.destroy_process_group()
if rank == 0:
    print("DYN_CHECK_OK" if res.item() == 1.0 else "DYN_CHECK_FAIL", flush=True)
