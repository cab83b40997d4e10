"""one GPU: MatMult and CG iteration time on an nx x ny x nz brick (default 512 x 512 x 64: what one rank of the 8-GPU
run holds, without the exchange) under the switches of the environment (PBX_L2_HINTS, PBX_UPDR_YFRONT ...)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import poissbox_b200 as pbx

nx, ny, nz = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 512, 64)
its = int(sys.argv[4]) if len(sys.argv) > 4 else 200
dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
h = pbx.Handle(nx, ny, nz, dx)
h.use_current_stream()
f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda") * 2 - 1
out = h.empty()


def tm(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


t_l = tm(lambda: h.lapl(f, out), 50)
t_ld = tm(lambda: h.lapl_dot(f, out), 50)
b = h.lapl(f)
res = []
t_cg = tm(lambda: res.append(h.cg_solve(b, rtol=1e-30, maxit=its)[1]), 3) / its
pm = h.lapl_profile(f, out, reps=10)
env = {k: v for k, v in os.environ.items() if k.startswith("PBX_")}
print(f"{nx}x{ny}x{nz}: lapl {t_l:.1f} us, lapl+dot {t_ld:.1f} us, passes alone x {pm[0] * 1e3:.1f} y {pm[1] * 1e3:.1f} z {pm[2] * 1e3:.1f}; "
      f"CG iteration {t_cg:.1f} us ({res[-1]} its per solve)  {env}", flush=True)
h.close()
