"""torchrun worker: checks the NCCL z-slab path (pbx_create with an ncclComm_t) on N GPUs against a
single-GPU evaluation of the same global problem, for the Laplacian, grad / div / interp and the CG.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29555 tools/dist_check.py [n]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import poissbox_b200 as pbx

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
NZ = int(sys.argv[2]) if len(sys.argv) > 2 else n      # optional: n x n x NZ brick (thin slabs on few GPUs)
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
g = torch.Generator(device=dev).manual_seed(1234)          # same seed: every rank builds the global field
f = torch.rand((NZ, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
whole = pbx.Handle(n, n, NZ, dx, device=local)
ref = whole.lapl(f)
h = pbx.Handle(n, n, nzl, dx, device=local, comm=comm.value)
mine = f[rank * nzl:(rank + 1) * nzl].contiguous()
out = h.lapl(mine)
torch.cuda.synchronize()
scale = ref.abs().max().item()
err = (out - ref[rank * nzl:(rank + 1) * nzl]).abs().max().item() / scale
w, dot = h.lapl_dot(mine)
dref = torch.dot(f.flatten(), ref.flatten()).item()
derr = abs(dot.item() - dref) / abs(dref)

# grad / div / interp over the communicator against the whole brick (FAST line operators)
v = torch.rand((3, NZ, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
sl = slice(rank * nzl, (rank + 1) * nzl)
gerr = 0.0
for name, got, want in (("grad", h.grad(mine), whole.grad(f)[:, sl]),
                        ("div", h.div(v[:, sl].contiguous()), whole.div(v)[sl]),
                        ("interp", h.interp(mine), whole.interp(f)[sl]),
                        ("interp_div", h.interp(mine, +1), whole.interp(f, +1)[sl])):
    torch.cuda.synchronize()
    gerr = max(gerr, (got - want).abs().max().item() / want.abs().max().item())
# the 2nd-order star on slabs (one plane each way): bit-identical to the whole brick
serr = 0.0 if torch.equal(h.star(mine), whole.star(f)[sl]) else 1.0
gerr = max(gerr, serr)

# CG: b = A x_true on the global grid; the slab solve must take the same iterations (+-1)
b = ref
CGMAX = int(os.environ.get("PBX_CHECK_CG_MAXIT", "2000"))
x1, its1, rn1, why1, hist1 = whole.cg_solve(b, rtol=1e-8, maxit=CGMAX)
bl = b[rank * nzl:(rank + 1) * nzl].contiguous()
xs, its2, rn2, why2, hist2 = h.cg_solve(bl, rtol=1e-8, maxit=CGMAX)
torch.cuda.synchronize()
xerr = (xs - x1[rank * nzl:(rank + 1) * nzl]).norm().item() / x1.norm().item()
m = min(len(hist1), len(hist2)) // 2
herr = float(np.max(np.abs(hist1[:m] - hist2[:m]) / hist1[:m]))
# the multigrid-preconditioned CG on slabs (PBX_CHECK_MG=0 skips it): same cycle, same iteration count
mg_note = "skipped"
mg_ok = True
if os.environ.get("PBX_CHECK_MG", "1") == "1":
    from poissbox_b200 import _lib

    hh = 2 * np.pi / n
    c = (torch.arange(n, dtype=torch.float64, device=dev) + 0.5) * hh
    cz_ = (torch.arange(NZ, dtype=torch.float64, device=dev) + 0.5) * (2 * np.pi / NZ)
    u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(cz_)[:, None, None]).contiguous()
    wm = pbx.Handle(n, n, NZ, (hh, hh, 2 * np.pi / NZ), device=local)
    hm = pbx.Handle(n, n, nzl, (hh, hh, 2 * np.pi / NZ), device=local, comm=comm.value)
    bu = wm.lapl(u)
    wm.set_pc(_lib.PC_MG, 2)
    hm.set_pc(_lib.PC_MG, 2)
    zw = wm.pc_apply(bu)
    zs = hm.pc_apply(bu[sl].contiguous())
    torch.cuda.synchronize()
    zerr = (zs - zw[sl]).abs().max().item() / zw.abs().max().item()
    xw, itw, _, whyw, _ = wm.cg_solve(bu, rtol=1e-8, maxit=100)
    xq, itq, _, whyq, _ = hm.cg_solve(bu[sl].contiguous(), rtol=1e-8, maxit=100)
    torch.cuda.synchronize()
    xmerr = (xq - xw[sl]).norm().item() / xw.norm().item()
    mg_ok = zerr <= 1e-13 and whyw == whyq == 2 and abs(itw - itq) <= 1 and xmerr <= 1e-6
    mg_note = f"vcycle err {zerr:.2e} pcg its {itw} vs {itq} reasons {whyw},{whyq} xerr {xmerr:.2e}"
    hm.close()
    wm.close()
ok = mg_ok and err <= 1e-13 and gerr <= 1e-13 and derr <= 1e-12 and why1 == why2 and abs(its1 - its2) <= 1 and xerr <= 1e-5 and herr <= 1e-6
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
print(f"rank {rank}/{world}: lapl err {err:.2e} grad/div/interp err {gerr:.2e} dot err {derr:.2e} cg its {its1} vs {its2} reasons {why1},{why2} "
      f"xerr {xerr:.2e} hist err {herr:.2e}; multigrid: {mg_note} -> {'OK' if ok else 'FAIL'}", flush=True)
h.close()
whole.close()
pbx.LIB.pbx_comm_destroy(comm)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK_OK" if res.item() == 1.0 else "DIST_CHECK_FAIL", flush=True)
sys.exit(0 if res.item() == 1.0 else 1)
