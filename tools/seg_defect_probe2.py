"""one GPU: regression check of the segmented TMA z tiles -- the z pass with its fused dot, many applies, against the
generic kernels (same arithmetic, same bits expected).  The output field is pre-filled with NaN, so a store that never
happened would show as NaN.  History: until round 2 the tile buffers were released to the TMA refill without
fence.proxy.async and 18 of 150 applies of (48, 512, 1088) returned one wrong 64-row box
(profiles/r2_seg_defect_rootcause.log, profiles/r2_seg_defect_fix_confirm.log: PBX_YZ_DBG 3 = unfenced, 0 = fenced; the
switch existed for those two runs only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import poissbox_b200 as pbx

shape = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (48, 512, 1088)
reps = int(os.environ.get("REPS", "40"))
variants = [0]
nx, ny, nz = shape
dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
g = torch.Generator(device="cuda").manual_seed(7)
f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
os.environ["PBX_NO_TMA"] = "1"
hg = pbx.Handle(nx, ny, nz, dx)
os.environ.pop("PBX_NO_TMA")
ref = hg.lapl(f)
hg.close()
out = torch.empty_like(ref)
for v in variants:
    h = pbx.Handle(nx, ny, nz, dx)
    fails, dumped, t = 0, False, 0.0
    for rep in range(reps):
        out.fill_(float("nan"))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.lapl_dot(f, out)
        e1.record()
        torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
        bad = out != ref
        nbad = int(bad.sum().item())
        if nbad:
            fails += 1
            if not dumped:
                dumped = True
                nan = int(torch.isnan(out).sum().item())
                idx = bad.nonzero()[0].tolist()
                z0, y0, x0 = idx
                lo, lr = out[:, y0, x0].cpu().numpy(), ref[:, y0, x0].cpu().numpy()
                zb = np.nonzero(lo != lr)[0]
                print(f"  variant {v} rep {rep}: {nbad} values differ, {nan} NaN; first bad line (x {x0}, y {y0}): bad z "
                      f"{zb.min()}..{zb.max()} ({len(zb)} planes)", flush=True)
                with np.printoptions(precision=6, linewidth=200):
                    k = zb.min()
                    print("   out", lo[max(0, k - 4):k + 12])
                    print("   ref", lr[max(0, k - 4):k + 12])
                    d = np.abs(lo - lr)
                    print("   |diff| by chunk", np.array([d[c * 16:(c + 1) * 16].max() for c in range(nz // 16)]))
    print(f"{shape} {fails} of {reps} lapl_dot runs differ from the generic kernels; {t / reps:.3f} ms per apply",
          flush=True)
    h.close()
