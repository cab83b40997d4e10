"""times grad / div / interp (both schedules) and the batched general-coefficient tridiagonal solves"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poissbox_b200 as pbx
from poissbox_b200 import LIB, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = pbx.Handle(n, n, n, (1.0 / n,) * 3); h.use_current_stream()
f = torch.rand((n, n, n), dtype=torch.float64, device="cuda") * 2 - 1
v = torch.rand((3, n, n, n), dtype=torch.float64, device="cuda") * 2 - 1
g3, s1 = h.empty(3), h.empty()
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
N = n ** 3
for mode, name in ((pbx.MODE_FAST, "FAST"), (pbx.MODE_REFERENCE, "REFERENCE")):
    h.mode = mode
    tg, td, ti, tl = tm(lambda: h.grad(f, g3)), tm(lambda: h.div(v, s1)), tm(lambda: h.interp(f, -1, s1)), tm(lambda: h.lapl(f, s1))
    print(f"{name:9s} {n}^3: grad {tg:.3f} ms ({N/tg/1e6:.1f} GDoF/s)  div {td:.3f} ms ({N/td/1e6:.1f})  interp {ti:.3f} ms ({N/ti/1e6:.1f})  lapl {tl:.3f} ms ({N/tl/1e6:.1f})")
# FAST grad / div / interp through the TMA-pipelined line operators (PBX_LINEOP_TMA=1, read per call)
os.environ["PBX_LINEOP_TMA"] = "1"
h.mode = pbx.MODE_FAST
tg, td, ti = tm(lambda: h.grad(f, g3)), tm(lambda: h.div(v, s1)), tm(lambda: h.interp(f, -1, s1))
print(f"FAST+TMA  {n}^3: grad {tg:.3f} ms ({N/tg/1e6:.1f} GDoF/s)  div {td:.3f} ms ({N/td/1e6:.1f})  interp {ti:.3f} ms ({N/ti/1e6:.1f})")
os.environ.pop("PBX_LINEOP_TMA")
# batched general-coefficient tridiagonal solves: n-point lines, element-major layout (coalesced)
for ln in (64, 512, 2048):
    nl = (1 << 24) // ln
    a, b, c, d = (torch.rand((ln, nl), dtype=torch.float64, device="cuda") for _ in range(4))
    b += 2.5
    ptr = [ctypes.c_void_p(t.data_ptr()) for t in (a, b, c, d)]
    tp = tm(lambda: check(LIB.pbx_tdma_periodic_batch_device(ln, nl, nl, 1, *ptr, None)))
    b2 = b.clone()
    ptr2 = [ctypes.c_void_p(t.data_ptr()) for t in (a, b2, c, d)]
    def run_tdma():
        b2.copy_(b)
        check(LIB.pbx_tdma_batch_device(ln, nl, nl, 1, *ptr2, None))
    tt = tm(run_tdma)
    pts = ln * nl
    print(f"tdma batch n={ln} lines={nl}: tdma_periodic {tp:.3f} ms ({pts/tp/1e6:.2f} Gpt/s, {48*pts/tp/1e6:.0f} GB/s alg)  tdma(+copy) {tt:.3f} ms ({pts/tt/1e6:.2f} Gpt/s)")

# line-major batches (contiguous lines): the generic thread-per-line kernels against the TMA-tile
# kernels of pbx_tdma_tma.cu (PBX_TDMA_TMA=1, read per call)
for ln, nl in ((64, 262144), (512, 32768), (2048, 8192)):
    a, b, c, d = (torch.rand((nl, ln), dtype=torch.float64, device="cuda") for _ in range(4))
    b += 2.5
    b2 = b.clone()
    ptr = [ctypes.c_void_p(t.data_ptr()) for t in (a, b, c, d)]
    ptr2 = [ctypes.c_void_p(t.data_ptr()) for t in (a, b2, c, d)]

    def run_tdma():
        b2.copy_(b)
        check(LIB.pbx_tdma_batch_device(ln, nl, 1, ln, *ptr2, None))

    row = []
    for tma in ("0", "1"):
        os.environ["PBX_TDMA_TMA"] = tma
        tp = tm(lambda: check(LIB.pbx_tdma_periodic_batch_device(ln, nl, 1, ln, *ptr, None)))
        tt = tm(run_tdma)
        row.append(f"TMA={tma}: periodic {ln * nl / tp / 1e6:.2f} Gpt/s, tdma(+copy) {ln * nl / tt / 1e6:.2f} Gpt/s")
    os.environ.pop("PBX_TDMA_TMA")
    print(f"tdma LINE-MAJOR n={ln} lines={nl}: " + "   ".join(row))
