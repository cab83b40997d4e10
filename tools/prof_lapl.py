"""Profiling driver: a few FAST compact-Laplacian applies (and optionally CG iterations) on an n^3
field.  Used under ncu (profiles/README.md); prints per-pass CUDA-event timings when run plainly."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import poissbox_b200 as pbx

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--shape", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"), help="brick instead of n^3")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--cg-its", type=int, default=0)
a = ap.parse_args()
n = a.n
nx, ny, nz = a.shape if a.shape else (n, n, n)
h = pbx.Handle(nx, ny, nz, (1.0 / nx, 1.0 / ny, 1.0 / nz))
h.use_current_stream()
g = torch.Generator(device="cuda").manual_seed(1234)
f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
out = h.empty()
for _ in range(a.reps):
    h.lapl(f, out)
torch.cuda.synchronize()
if a.cg_its:
    b = h.lapl(f)
    h.cg_solve(b, rtol=1e-30, maxit=a.cg_its)
    torch.cuda.synchronize()
ms = h.lapl_profile(f, out, reps=5)
print("pass ms (x,y,z):", ms, "total", sum(ms), "GDoF/s", nx * ny * nz / sum(ms) / 1e6)
