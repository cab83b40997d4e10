"""times the phases of one slab of the z-decomposed Laplacian on one GPU (no exchange)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poissbox_b200 as pbx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nzl = n // P
dx = (1.0 / n,) * 3
h = pbx.Handle(n, n, nzl, dx, slab=(0, P)); h.use_current_stream()
w = pbx.Handle(n, n, nzl, dx); w.use_current_stream()
f = torch.rand((nzl, n, n), dtype=torch.float64, device="cuda") * 2 - 1
out = h.empty()
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_whole = tm(lambda: w.lapl(f, out))
t1 = tm(lambda: h.slab_phase1(f))
t2 = tm(lambda: h.slab_phase2(out))
pm = w.lapl_profile(f, out, reps=5)
print(f"n={n} P={P} nzl={nzl}: periodic brick lapl {t_whole:.3f} ms (x,y,z = {pm[0]:.3f},{pm[1]:.3f},{pm[2]:.3f}); "
      f"slab phase1 (x+y+moments) {t1:.3f} ms -> moments ~{t1 - pm[0] - pm[1]:.3f}; phase2 (open z + corrections) {t2:.3f} ms vs z {pm[2]:.3f}")
