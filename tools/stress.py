"""Stress: repeated CG solves / Laplacian applies must be bit-reproducible run to run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poissbox_b200 as pbx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ncg = int(sys.argv[2]) if len(sys.argv) > 2 else 6
hh = 2 * np.pi / n
c = (torch.arange(n, dtype=torch.float64, device="cuda") + 0.5) * hh
ref_hist = None
for rep in range(ncg):
    u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
    h = pbx.Handle(n, n, n, (hh,) * 3); h.use_current_stream()
    b = h.lapl(u)
    x = h.empty()
    del u
    torch.cuda.synchronize()
    bsum = b.double().sum().item(); babs = b.abs().sum().item()
    x, its, rnorm, reason, hist = h.cg_solve(b, x, rtol=1e-8, maxit=20000 if rep % 2 == 0 else 3000)
    torch.cuda.synchronize()
    same = ref_hist is None or (len(hist) == len(ref_hist) and np.array_equal(hist, ref_hist))
    if ref_hist is None: ref_hist = hist.copy()
    firstdiff = -1
    if not same:
        m = min(len(hist), len(ref_hist)); d = np.nonzero(hist[:m] != ref_hist[:m])[0]
        firstdiff = int(d[0]) if len(d) else m
    print(f"rep {rep}: its {its} reason {reason} rel {rnorm/hist[0]:.3e} bsum {bsum:.17e} babs {babs:.17e} same {same} firstdiff {firstdiff}", flush=True)
    h.close()
