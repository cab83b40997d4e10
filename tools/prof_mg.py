"""times the 2nd-order star, one multigrid V(nu,nu) cycle (nu = 1, 2, 3) and the preconditioned CG
on an n^3 box (default 512): CUDA events around repeated calls through the C ABI"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import poissbox_b200 as pbx
from poissbox_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hh = 2 * np.pi / n
h = pbx.Handle(n, n, n, (hh,) * 3)
h.use_current_stream()
c = (torch.arange(n, dtype=torch.float64, device="cuda") + 0.5) * hh
u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
out = h.empty()


def tm(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


N = n**3
t = tm(lambda: h.star(u, out))
print(f"star apply {n}^3: {t:.3f} ms = {N / t / 1e6:.1f} GDoF/s, {16 * N / t / 1e6:.0f} GB/s at 16 B/DoF")
t = tm(lambda: h.lapl(u, out))
print(f"compact lapl: {t:.3f} ms")
ts = {}
for nu in (1, 2, 3):
    h.set_pc(_lib.PC_MG, nu)
    ts[nu] = tm(lambda: h.pc_apply(u, out))
    print(f"pc_apply V({nu},{nu}) (incl. two mean reductions): {ts[nu]:.3f} ms")
print(f"  per fine-level Jacobi sweep ~ {(ts[3] - ts[1]) / 4 * 7 / 8:.3f} ms ({24 * N / ((ts[3] - ts[1]) / 4 * 7 / 8) / 1e6:.0f} GB/s at 24 B/DoF)")
b = h.lapl(u)
for nu in (1, 2):
    h.set_pc(_lib.PC_MG, nu)
    h.cg_solve(b, rtol=1e-8, maxit=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x, its, rn, why, hist = h.cg_solve(b, rtol=1e-8)
    e1.record()
    torch.cuda.synchronize()
    err = ((x - x.mean()) - (u - u.mean())).norm().item() / (u - u.mean()).norm().item()
    print(f"PCG V({nu},{nu}): {its} its, reason {why}, {e0.elapsed_time(e1):.1f} ms, solution error {err:.2e}")
h.operator = 1
bs = h.star(u)
h.set_pc(_lib.PC_MG, 2)
x, its, rn, why, hist = h.cg_solve(bs, rtol=1e-8)
print(f"PCG on the star operator itself: {its} its, reason {why}")
h.close()
