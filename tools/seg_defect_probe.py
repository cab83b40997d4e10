"""one GPU: where the segmented TMA y / z tiles (PBX_TMA_SEG=1, off by default) go wrong -- error of the TMA Laplacian
against the generic kernels' (same arithmetic, same bits expected) by y chunk and by z chunk, for bricks with a
segmented y line, a segmented z line, or both"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import poissbox_b200 as pbx

os.environ["PBX_TMA_SEG"] = "1"
shapes = [(48, 640, 1088), (48, 640, 512), (48, 512, 1088)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for nx, ny, nz in shapes:
    dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
    g = torch.Generator(device="cuda").manual_seed(7)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    os.environ["PBX_NO_TMA"] = "1"
    hg = pbx.Handle(nx, ny, nz, dx)
    os.environ.pop("PBX_NO_TMA")
    ref = hg.lapl(f)
    hg.close()
    h = pbx.Handle(nx, ny, nz, dx)
    for rep in range(int(os.environ.get("REPS", "8"))):
        if rep < 2:
            out = h.lapl(f)
        else:
            out, _ = h.lapl_dot(f)     # the z pass with its fused dot; x pass walks back to front, y pass front to back
        torch.cuda.synchronize()
        bad = (out != ref)
        nbad = int(bad.sum().item())
        line = f"{(nx, ny, nz)} rep {rep} ({'lapl' if rep < 2 else 'lapl_dot'}): {nbad} of {out.numel()} values differ"
        if nbad:
            zc = bad.reshape(nz // 16, 16, ny, nx).any(dim=1).any(dim=1).any(dim=1).nonzero().flatten().tolist()
            yc = bad.reshape(nz, ny // 16, 16, nx).any(dim=2).any(dim=0).any(dim=1).nonzero().flatten().tolist()
            xs = bad.any(dim=0).any(dim=0).nonzero().flatten().tolist()
            planes = bad.any(dim=1).any(dim=1).nonzero().flatten().tolist()
            line += (f"; z chunks {zc[:40]}{'...' if len(zc) > 40 else ''} ({len(zc)}); y chunks {yc[:40]}{'...' if len(yc) > 40 else ''} ({len(yc)}); "
                     f"x {xs[:8]}..{xs[-3:]} ({len(xs)}); planes {len(planes)}; max rel diff "
                     f"{((out - ref).abs().max() / ref.abs().max()).item():.2e}")
        print(line, flush=True)
    h.close()
