"""Diagnostic: does stale device memory influence the CG?  Fill freshly freed memory with NaN before
every handle creation / solve and compare residual histories between repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poissbox_b200 as pbx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
hh = 2 * np.pi / n
c = (torch.arange(n, dtype=torch.float64, device="cuda") + 0.5) * hh
ref = None
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for rep in range(nrep):
    if rep > 0:
        junk = [torch.full((1 << 27,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(gb)]
        torch.cuda.synchronize(); del junk; torch.cuda.empty_cache()
    u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
    h = pbx.Handle(n, n, n, (hh,) * 3); h.use_current_stream()
    b = h.lapl(u); x = h.empty(); del u
    x, its, rn, why, hist = h.cg_solve(b, x, rtol=1e-8, maxit=1500)
    torch.cuda.synchronize()
    if ref is None: ref = hist.copy()
    m = min(len(hist), len(ref)); d = np.nonzero(hist[:m] != ref[:m])[0]
    print(f"rep {rep}: its {its} reason {why} rel {rn/hist[0]:.3e} nan_in_hist {int(np.isnan(hist).sum())} first_diff {int(d[0]) if len(d) else -1}", flush=True)
    if len(d): print("   hist around:", hist[max(0,d[0]-1):d[0]+3], "ref:", ref[max(0,d[0]-1):d[0]+3])
    h.close(); del b, x
