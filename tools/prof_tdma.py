"""the batched general-coefficient tridiagonal solves alone (tdma, tdma_periodic; both layouts), a few calls each:
run plainly for CUDA-event timings, under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
for the kernels' own durations and DRAM traffic (48 B/point algorithmic)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poissbox_b200 as pbx
from poissbox_b200 import LIB, check

reps = int(os.environ.get("REPS", "5"))
lines = [int(v) for v in sys.argv[1:]] or [64, 512, 2048]
for L in lines:
    nl = (1 << 24) // L
    for layout in ("elem_major", "line_major"):
        shape = (L, nl) if layout == "elem_major" else (nl, L)
        es, ls = (nl, 1) if layout == "elem_major" else (1, L)
        A, B, C, D = (torch.rand(shape, dtype=torch.float64, device="cuda") for _ in range(4))
        B += 2.5
        ptr = [ctypes.c_void_p(t.data_ptr()) for t in (A, B, C, D)]
        fn = lambda: check(LIB.pbx_tdma_periodic_batch_device(L, nl, es, ls, *ptr, None))
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        print(f"tdma_periodic n={L} lines={nl} {layout}: {t:.3f} ms  {L * nl / t / 1e6:.2f} Gpt/s  {48 * L * nl / t / 1e6:.0f} GB/s (48 B/pt)", flush=True)
        del A, B, C, D
