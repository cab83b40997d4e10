"""torchrun worker: time per CG iteration over NCCL for a fixed number of iterations"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import poissbox_b200 as pbx
from poissbox_b200 import LIB, check
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
its = int(sys.argv[2]) if len(sys.argv) > 2 else 200
nzl = n // world
idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    raw = (ctypes.c_ubyte * 128)(); check(LIB.pbx_comm_unique_id(raw))
    idbuf = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
dist.broadcast(idbuf, 0)
raw = (ctypes.c_ubyte * 128)(*idbuf.cpu().tolist()); comm = ctypes.c_void_p()
check(LIB.pbx_comm_init_rank(raw, world, rank, local, ctypes.byref(comm)))
h = pbx.Handle(n, n, nzl, (1.0 / n,) * 3, device=local, comm=comm.value); h.use_current_stream()
g = torch.Generator(device=dev).manual_seed(rank)
f = torch.rand((nzl, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
b = h.lapl(f)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, it, rn, why, hist = h.cg_solve(b, rtol=1e-30, maxit=its)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0: print(f"world {world} n {n}: {it} its, {dt / max(it,1) * 1e3:.3f} ms/it  env {[k for k in os.environ if k.startswith('PBX_')]}", flush=True)
h.close(); LIB.pbx_comm_destroy(comm); dist.destroy_process_group()
