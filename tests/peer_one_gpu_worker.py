"""worker of tests/test_zslab_gpu.py::test_peer_boards_one_gpu: P slab handles of ONE process on ONE GPU, each on its
own stream and host thread, linked by pbx_slab_link_peers (boundary messages by direct stores, flag barrier, all-reduce
inside the CG's reduction kernels), against one handle on the whole brick.  Runs in a process of its own: a rank's
kernels spin on the device until the other ranks' kernels have run, so the ranks' streams must not share a hardware
queue (CUDA_DEVICE_MAX_CONNECTIONS is raised by the test), and a spin that never ends traps the context -- which
must not be the test session's."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poissbox_b200 as pbx
from poissbox_b200 import _lib


def main(P):
    import threading

    import torch

    nx, ny, nzl = 64, 32, 64
    nz = nzl * P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    g = torch.Generator(device="cuda").manual_seed(11)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    whole = pbx.Handle(nx, ny, nz, dx)
    ref = whole.lapl(f)
    x1, its1, _, why1, hist1 = whole.cg_solve(ref, rtol=1e-6, maxit=2000)
    whole.set_pc(_lib.PC_MG, 2)
    xm1, itm1, _, whym1, _ = whole.cg_solve(ref, rtol=1e-6, maxit=200)
    torch.cuda.synchronize()
    slabs = [pbx.Handle(nx, ny, nzl, dx, slab=(r, P)) for r in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    for h, s in zip(slabs, streams):
        h.set_stream(s.cuda_stream)
    pbx.Handle.slab_link_local(slabs)
    res = [None] * P

    def work(r):
        h = slabs[r]
        part = f[r * nzl:(r + 1) * nzl].contiguous()
        bpart = ref[r * nzl:(r + 1) * nzl].contiguous()
        torch.cuda.synchronize()
        with torch.cuda.stream(streams[r]):
            outs = [h.lapl(part) for _ in range(3)]
            x, its, _, why, hist = h.cg_solve(bpart, rtol=1e-6, maxit=2000)
            # the multigrid-preconditioned CG on slabs (halo exchanges per level, coarse levels gathered)
            h.set_pc(_lib.PC_MG, 2)
            xm, itm, _, whym, _ = h.cg_solve(bpart, rtol=1e-6, maxit=200)
            h.synchronize()
        res[r] = (outs, x, its, why, hist, xm, itm, whym)

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert all(r is not None for r in res), "a rank did not finish"
    scale = ref.abs().max().item()
    for r in range(P):
        outs, x, its, why, hist, xm, itm, whym = res[r]
        assert whym == whym1 and abs(itm - itm1) <= 1, (itm, itm1, whym, whym1)
        for o in outs:
            assert (o - ref[r * nzl:(r + 1) * nzl]).abs().max().item() <= 1e-13 * scale
        assert why == why1 == 2 and abs(its - its1) <= 1, (its, its1, why)
        assert (x - x1[r * nzl:(r + 1) * nzl]).abs().max().item() <= 1e-6 * x1.abs().max().item()
    for h in slabs + [whole]:
        h.close()
    print("PEER_ONE_GPU_OK", flush=True)


if __name__ == "__main__":
    try:
        main(int(sys.argv[1]))
    except BaseException:
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)   # a rank stuck on the device must not keep the interpreter from exiting
    os._exit(0)
