import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poissbox_b200 as pbx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hh = 2 * np.pi / n
def mk(no_tma, yz):
    os.environ["PBX_NO_TMA"] = no_tma; os.environ["PBX_TMA_YZ"] = yz
    h = pbx.Handle(n, n, n, (hh,) * 3); h.use_current_stream()
    os.environ.pop("PBX_NO_TMA"); os.environ.pop("PBX_TMA_YZ")
    return h
hT, hG = mk("0", "1"), mk("1", "0")
g = torch.Generator(device="cuda").manual_seed(1)
f = torch.rand((n, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
bad = 0
wG, dG = hG.lapl_dot(f)
for i in range(reps):
    wT, dT = hT.lapl_dot(f)
    if not torch.equal(wT, wG) or not torch.equal(dT, dG):
        bad += 1
        d = (wT != wG).nonzero()
        print("mismatch at rep", i, "count", d.shape[0], "first", d[:3].tolist(), "dot", dT.item(), dG.item())
        if bad > 3: break
print("stress: bad =", bad, "of", reps)
# CG with both
c = (torch.arange(n, dtype=torch.float64, device="cuda") + 0.5) * hh
u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
for name, h in (("generic", hG), ("tma", hT)):
    b = h.lapl(u)
    t0 = time.perf_counter()
    x, its, rnorm, reason, hist = h.cg_solve(b, rtol=1e-8, maxit=3000)
    torch.cuda.synchronize()
    print(name, "its", its, "reason", reason, "rel", rnorm / hist[0], "time", time.perf_counter() - t0)
    k = [0, 1, 2, 5, 10, 50, 100, 200, 400, 800, 1000, 1100]
    print("  hist", [(i, float(hist[i] / hist[0])) for i in k if i < len(hist)])
