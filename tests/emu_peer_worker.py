"""worker of test_zslab_cpu.py::test_peer_boards_world*: one process per rank; the ranks share their
receive buffers (POSIX shared memory standing in for NVLink peer mappings), link them with
pbx_slab_link_peers and from then on run the library's own multi-rank code paths on the CPU
kernel-logic harness (tests/emu) WITHOUT any host-side exchange: boundary sweeps storing into the
neighbours' buffers, the flag barrier, the all-reduce inside the CG's reduction kernel.  gloo is
only the rendezvous (and the referee of the final verdict)."""
import ctypes
import os
import sys
from multiprocessing import resource_tracker, shared_memory

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import emu_lib
from poissbox_b200 import _lib

rank, world, tag = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dist.init_process_group("gloo", rank=rank, world_size=world)
nx, ny, nzl = 16, 16, 64
nz = nzl * world
dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
rng = np.random.default_rng(4242)
f = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
v = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz, 3)))
mine = slice(rank * nzl, (rank + 1) * nzl)
h = emu_lib.EmuHandle(nx, ny, nzl, dx, slab=(rank, world))
lib = h.lib

# every rank creates its own (zero-filled) segment, then maps everybody's
nbytes = ctypes.c_size_t()
emu_lib.check(lib, lib.pbx_slab_recv_bytes(h._h, ctypes.byref(nbytes)))
own = shared_memory.SharedMemory(name=f"pbxpeer_{tag}_{rank}", create=True, size=nbytes.value)
dist.barrier()
segs = [own if r == rank else shared_memory.SharedMemory(name=f"pbxpeer_{tag}_{r}") for r in range(world)]
for r, sgm in enumerate(segs):
    if r != rank:   # only the creator unlinks (Python < 3.13 registers attachments as well)
        resource_tracker.unregister(sgm._name, "shared_memory")
addr = [ctypes.addressof(ctypes.c_char.from_buffer(s.buf)) for s in segs]
bufs = (ctypes.c_void_p * world)(*addr)
emu_lib.check(lib, lib.pbx_slab_link_peers(h._h, bufs, world))
dist.barrier()

errs = {}
whole = emu_lib.EmuHandle(nx, ny, nz, dx)
ref = whole.lapl(f)
lib.pbx_launch_count.restype = ctypes.c_longlong
for rep in range(3):   # three rounds: both parities of the receive arrays and a reused one
    l0 = lib.pbx_launch_count(h._h)
    out = h.lapl(np.asfortranarray(f[:, :, mine]))
    errs[f"lapl{rep}"] = np.max(np.abs(out - ref[:, :, mine])) / np.max(np.abs(ref))
    assert lib.pbx_launch_count(h._h) - l0 == 5   # x, y, boundary sweep, barrier, z
for name, fn, src, want in (("grad", h.grad, f, whole.grad(f)), ("div", h.div, v, whole.div(v)),
                            ("interp", h.interp, f, whole.interp(f)), ("star", h.star, f, whole.star(f))):
    got = fn(np.asfortranarray(src[:, :, mine]))
    errs[name] = np.max(np.abs(got - want[:, :, mine])) / np.max(np.abs(want))
# the fused z pass + dot + all-reduce
outd, dot = h.lapl_dot(np.asfortranarray(f[:, :, mine]))
errs["lapl_of_dot"] = np.max(np.abs(outd - ref[:, :, mine])) / np.max(np.abs(ref))   # the in-CG tile order and sweep direction
ref_dot = float(np.vdot(f, ref))
errs["dot"] = abs(dot - ref_dot) / abs(ref_dot)
# the stand-alone all-reduce: every rank must see the same bits
vals = np.array([rank + 1.0, 0.1 * (rank + 1), -3.0, 1e-3 * rank])
emu_lib.check(lib, lib.pbx_allreduce_sum(h._h, emu_lib.ptr(vals), 4))
want_vals = [sum(r + 1.0 for r in range(world)), sum(0.1 * (r + 1) for r in range(world)), -3.0 * world,
             sum(1e-3 * r for r in range(world))]
errs["allreduce"] = float(np.max(np.abs(vals - want_vals)))
gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(vals.copy()))
same_bits = all(torch.equal(g, gathered[0]) for g in gathered)

MAXIT = 30   # bounded: the harness runs ~10 iterations a second
# the distributed CG against the single-handle CG on the whole brick: same iteration count, same
# residual history (to rounding of the differently associated sums), same solution
# (b = A f with a full-spectrum f: for smooth right-hand sides the history is governed by rounding
# noise -- SURVEY section 7 -- and differently associated sums part ways after a few dozen iterations)
b = ref
x1, its1, rn1, why1, hist1 = whole.cg_solve(b, rtol=1e-8, maxit=MAXIT)
xs, its, rn, why, hist = h.cg_solve(np.asfortranarray(b[:, :, mine]), rtol=1e-8, maxit=MAXIT)
errs["cg_x"] = np.max(np.abs(xs - x1[:, :, mine])) / np.max(np.abs(x1))
n = min(len(hist), len(hist1))
errs["cg_hist"] = float(np.max(np.abs(hist[:n] - hist1[:n]) / hist1[:n]))
cg_ok = why == why1 and its == its1
# a solve that converges: the iteration issued after convergence is a no-op on every rank (the
# guarded kernels skip their exchange everywhere or nowhere), and the boards stay in step after it
_, its2a, _, why2a, _ = whole.cg_solve(b, rtol=3e-1, maxit=MAXIT)
_, its2, _, why2, _ = h.cg_solve(np.asfortranarray(b[:, :, mine]), rtol=3e-1, maxit=MAXIT)
cg_ok = cg_ok and why2 == why2a == 2 and its2 == its2a
out = h.lapl(np.asfortranarray(f[:, :, mine]))
errs["lapl_after_cg"] = np.max(np.abs(out - ref[:, :, mine])) / np.max(np.abs(ref))
# the multigrid preconditioner on slabs (distributed levels with one-plane halo exchanges, the
# coarse levels all-gathered and solved redundantly) against the single-rank cycle: the V-cycle is
# the same arithmetic point for point, so z carries the same bits; the PCG the same iterations
c = [(np.arange(m) + 0.5) * 2 * np.pi / m for m in (nx, ny, nz)]
u = np.exp(np.sin(c[0])[:, None, None] + np.sin(c[1])[None, :, None] + np.sin(c[2])[None, None, :])
hp = (2 * np.pi / nz,) * 3   # isotropic spacing: point Jacobi smooths it well
whole_p = emu_lib.EmuHandle(nx, ny, nz, hp)
h_p = emu_lib.EmuHandle(nx, ny, nzl, hp, slab=(rank, world))
own_p = shared_memory.SharedMemory(name=f"pbxpeerp_{tag}_{rank}", create=True, size=nbytes.value)
dist.barrier()
segs_p = [own_p if r == rank else shared_memory.SharedMemory(name=f"pbxpeerp_{tag}_{r}") for r in range(world)]
for r, sgm in enumerate(segs_p):
    if r != rank:
        resource_tracker.unregister(sgm._name, "shared_memory")
bufs_p = (ctypes.c_void_p * world)(*[ctypes.addressof(ctypes.c_char.from_buffer(s_.buf)) for s_ in segs_p])
emu_lib.check(lib, lib.pbx_slab_link_peers(h_p._h, bufs_p, world))
dist.barrier()
bu = whole_p.lapl(np.asfortranarray(u))
whole_p.set_pc(_lib.PC_MG, 2)
h_p.set_pc(_lib.PC_MG, 2)
zw = whole_p.pc_apply(bu)
zs = h_p.pc_apply(np.asfortranarray(bu[:, :, mine]))
errs["mg_vcycle"] = float(np.max(np.abs(zs - zw[:, :, mine])) / np.max(np.abs(zw)))
xw, itw, _, whyw, histw = whole_p.cg_solve(bu, rtol=1e-8, maxit=40)
xq, itq, _, whyq, histq = h_p.cg_solve(np.asfortranarray(bu[:, :, mine]), rtol=1e-8, maxit=40)
errs["pcg_x"] = float(np.max(np.abs(xq - xw[:, :, mine])) / np.max(np.abs(xw)))
cg_ok = cg_ok and whyq == whyw == 2 and itq == itw and itq <= 20
for nu_ in (1, 3):   # one and three smoothing sweeps take other branches of the halo protocol
    whole_p.set_pc(_lib.PC_MG, nu_)
    h_p.set_pc(_lib.PC_MG, nu_)
    zw = whole_p.pc_apply(bu)
    zs = h_p.pc_apply(np.asfortranarray(bu[:, :, mine]))
    errs[f"mg_vcycle_nu{nu_}"] = float(np.max(np.abs(zs - zw[:, :, mine])) / np.max(np.abs(zw)))
h_p.close()
whole_p.close()
del bufs_p
for s_ in segs_p:
    try:
        s_.close()
    except BufferError:
        pass
dist.barrier()
try:
    own_p.unlink()
except FileNotFoundError:
    pass
# ranks agree on the iteration count (the status word derives from all-reduced sums)
its_all = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(its_all, torch.tensor([its], dtype=torch.int64))
cg_ok = cg_ok and all(int(t) == its for t in its_all)

tol = {"mg_vcycle": 1e-14, "mg_vcycle_nu1": 1e-14, "mg_vcycle_nu3": 1e-14, "pcg_x": 1e-9, "star": 0.0, "dot": 1e-12, "allreduce": 1e-15, "cg_x": 1e-6, "cg_hist": 1e-9}
ok = all(e <= tol.get(k, 1e-13) for k, e in errs.items()) and same_bits and cg_ok
dist.barrier()
h.close()
whole.close()
del bufs, addr
for s in segs:
    try:
        s.close()
    except BufferError:
        pass
dist.barrier()
try:
    own.unlink()
except FileNotFoundError:
    pass
dist.destroy_process_group()
print(("EMU_PEER_OK " if ok else "EMU_PEER_FAIL ") + f"its {its} vs {its1} why {why}; {its2} vs {its2a} why {why2}; pcg {itq} vs {itw} why {whyq}; bits {same_bits} " + str(errs))
sys.exit(0 if ok else 1)
