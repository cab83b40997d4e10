"""one GPU: the FAST Laplacian (+ fused dot) of a shape with segmented y / z lines, run many times per kernel family;
every run of a family must carry the same bits, and the families must agree (tests/test_parity_gpu.py failed ONCE on
(16, 640, 1088) in round 2: is a kernel family non-deterministic?)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import poissbox_b200 as pbx

shapes = [(16, 640, 1088), (32, 16, 2048), (64, 1024, 32), (512, 512, 32)]
if len(sys.argv) > 1:   # shapes as nx,ny,nz ...
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
REPS = int(os.environ.get("REPS", "30"))
for nx, ny, nz in shapes:
    dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
    g = torch.Generator(device="cuda").manual_seed(7)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    fam = {}
    for name, no_tma, tma_yz in (("tma", "0", "1"), ("generic", "1", "0"), ("tma-x", "0", "0")):
        os.environ["PBX_NO_TMA"], os.environ["PBX_TMA_YZ"] = no_tma, tma_yz
        h = pbx.Handle(nx, ny, nz, dx)
        os.environ.pop("PBX_NO_TMA"); os.environ.pop("PBX_TMA_YZ")
        first, bad, worst = None, 0, 0.0
        for rep in range(REPS):
            w, dot = h.lapl_dot(f)
            torch.cuda.synchronize()
            if first is None:
                first = (w.clone(), dot.clone())
            elif not (torch.equal(w, first[0]) and torch.equal(dot, first[1])):
                bad += 1
                worst = max(worst, (w - first[0]).abs().max().item() / first[0].abs().max().item())
        fam[name] = first
        print(f"{(nx, ny, nz)} {name:8s}: {bad} of {REPS - 1} repeats differ from the first run (worst field diff {worst:.2e})", flush=True)
        h.close()
    same = all(torch.equal(fam["tma"][0], v[0]) for v in fam.values())
    print(f"{(nx, ny, nz)} families agree on the field: {same}", flush=True)
