"""torchrun worker: times the pieces of the z-slab MatMult / CG iteration over NCCL"""
import ctypes, os, sys, time
if os.environ.get("PBX_P2P_CH"):
    os.environ["NCCL_MIN_P2P_NCHANNELS"]=os.environ["PBX_P2P_CH"]; os.environ["NCCL_MAX_P2P_NCHANNELS"]="32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import poissbox_b200 as pbx
from poissbox_b200 import LIB, check
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NZ = int(sys.argv[2]) if len(sys.argv) > 2 else n      # optional: n x n x NZ brick (thin slabs on few GPUs)
nzl = NZ // world
idbuf = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    raw = (ctypes.c_ubyte * 128)(); check(LIB.pbx_comm_unique_id(raw))
    idbuf = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
dist.broadcast(idbuf, 0)
raw = (ctypes.c_ubyte * 128)(*idbuf.cpu().tolist()); comm = ctypes.c_void_p()
check(LIB.pbx_comm_init_rank(raw, world, rank, local, ctypes.byref(comm)))
h = pbx.Handle(n, n, nzl, (1.0 / n,) * 3, device=local, comm=comm.value); h.use_current_stream()
f = torch.rand((nzl, n, n), dtype=torch.float64, device=dev) * 2 - 1
out = h.empty(); sc = torch.zeros(4, dtype=torch.float64, device=dev)
def tm(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6
res = {}
res["lapl"] = tm(lambda: h.lapl(f, out))
w1 = pbx.Handle(n, n, nzl, (1.0 / n,) * 3, device=local); w1.use_current_stream()
res["lapl, periodic brick of the same size (no exchange)"] = tm(lambda: w1.lapl(f, out))
w1.close()
if os.environ.get("PBX_PROF_PHASES", "1") == "1":   # the unfused pieces (they leave the epoch counters in step)
    res["phase1 (x, y, boundary sweep)"] = tm(lambda: h.slab_phase1(f))
    res["exchange"] = tm(lambda: check(LIB.pbx_slab_exchange(h._h)))
    res["phase2 (slab z pass)"] = tm(lambda: h.slab_phase2(out))
bq = h.lapl(f)
def cgits():
    h.cg_solve(bq, rtol=1e-30, maxit=50)
t_dev, t_host = tm(cgits, reps=4)
res["CG iteration (50-iteration solves)"] = (t_dev / 50, t_host / 50)
res["allreduce_pbx(2 doubles)"] = tm(lambda: check(LIB.pbx_allreduce_sum(h._h, ctypes.c_void_p(sc.data_ptr()), 2)))
res["allreduce_torch(2 doubles)"] = tm(lambda: dist.all_reduce(sc[:2]))
big = torch.zeros(2 * 1024 * 1024, dtype=torch.float64, device=dev)
res["torch p2p 16MiB ring"] = tm(lambda: [r.wait() for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, big, (rank + 1) % world), dist.P2POp(dist.irecv, out.view(-1)[:big.numel()], (rank - 1) % world)])], reps=20)
if rank == 0:
    for k, (dv, host) in res.items(): print(f"{k:32s} device {dv:9.1f} us   host {host:9.1f} us", flush=True)
h.close(); LIB.pbx_comm_destroy(comm); dist.destroy_process_group()
