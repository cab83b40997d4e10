"""GPU experiment: the z-blocked x+y schedule (PBX_XY_BLOCK, fast_xy in pbx_api.cu) at 512^3 --
time per Laplacian apply and per pass for several block sizes -- and the segmented long-line
kernels on bricks with 1024- and 2048-point y / z lines (BASELINE configs[4])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import poissbox_b200 as pbx


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(shape, zb, reps=20):
    nx, ny, nz = shape
    os.environ["PBX_XY_BLOCK"] = str(zb)
    h = pbx.Handle(nx, ny, nz, (1.0 / nx, 1.0 / ny, 1.0 / nz))
    os.environ.pop("PBX_XY_BLOCK")
    h.use_current_stream()
    g = torch.Generator(device="cuda").manual_seed(1234)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    out = h.empty()
    for _ in range(3):
        h.lapl(f, out)
    t = timed(lambda: h.lapl(f, out), reps)
    td = timed(lambda: h.lapl_dot(f, out), reps)
    ms = h.lapl_profile(f, out, reps=5)
    n = nx * ny * nz
    print(f"{shape} zb={zb:3d}: apply {t:.3f} ms = {n / t / 1e6:.1f} GDoF/s ({80 * n / t / 1e6 / 6551:.3f} of HBM peak at 80 B/DoF)"
          f"  with dot {td:.3f} ms  passes x {ms[0]:.3f} y {ms[1]:.3f} z {ms[2]:.3f}", flush=True)
    chk = out.clone()
    h.close()
    return chk


if __name__ == "__main__":
    ref = None
    for zb in (0, 6, 9, 14, 18, 23, 28, 37):
        o = run((512, 512, 512), zb)
        if ref is None:
            ref = o
        elif not torch.equal(o, ref):
            print("   MISMATCH against the unblocked result")
    del ref, o
    torch.cuda.empty_cache()
    for shape in ((512, 1024, 256), (512, 256, 1024), (256, 2048, 256), (256, 256, 2048), (1024, 1024, 128), (2048, 256, 256)):
        run(shape, 0, reps=10)
        run(shape, 16, reps=10)
