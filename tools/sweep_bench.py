"""BASELINE configs[4]: standalone batched tridsol + compact grad / div / Laplacian sweeps with line
lengths 64-2048 along x, y and z on bricks of 2^27 points (the size of 512^3).  One GPU.  Prints one
JSON line per measurement and a summary table; `--out FILE` also writes the lines to FILE.

Algorithmic bytes: Laplacian 80 B/DoF; grad and div 32 B/DoF floor (the FAST schedule
moves 8 operators x 16 B = 128 B/DoF); general-coefficient tdma 48 B/point.
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import poissbox_b200 as pbx
from poissbox_b200 import LIB, check

PEAK = 6551.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def tm(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def shapes(logn):
    """bricks of 2^logn points with a line of L = 64 .. 2048 along each axis in turn"""
    out = []
    for L in (64, 128, 256, 512, 1024, 2048):
        rest = logn - L.bit_length() + 1
        a, b = 1 << ((rest + 1) // 2), 1 << (rest // 2)
        out += [("x", L, (L, a, b)), ("y", L, (a, L, b)), ("z", L, (a, b, L))]
    return out


def run(logn=27, skip_ops=False, out=None):
    class A:
        pass

    a = A()
    a.logn, a.skip_ops, a.out = logn, skip_ops, out
    lines = []

    def emit(rec):
        lines.append(rec)
        print(json.dumps(rec), flush=True)

    if not a.skip_ops:
        for axis, L, (nx, ny, nz) in shapes(a.logn):
            N = nx * ny * nz
            h = pbx.Handle(nx, ny, nz, (1.0 / nx, 1.0 / ny, 1.0 / nz))
            h.use_current_stream()
            f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda") * 2 - 1
            o = h.empty()
            t = tm(lambda: h.lapl(f, o))
            px, py, pz = h.lapl_profile(f, o, reps=3)
            emit({"op": "lapl", "axis": axis, "line": L, "brick": [nx, ny, nz], "ms": t, "GDoF_s": N / t / 1e6,
                  "frac_hbm": 80 * N / t / 1e6 / PEAK,
                  "pass_ms": {"x": px, "y": py, "z": pz},
                  "pass_frac_hbm": {"x": 24 * N / px / 1e6 / PEAK, "y": 32 * N / py / 1e6 / PEAK,
                                    "z": 24 * N / pz / 1e6 / PEAK}})
            v, g3 = h.empty(3), h.empty(3)
            v.uniform_(-1, 1)
            tg, td = tm(lambda: h.grad(f, g3)), tm(lambda: h.div(v, o))
            # 32 B/DoF is the floor (1 field in, 3 out); the one-sweep-per-axis schedule moves 112 B/DoF
            emit({"op": "grad", "axis": axis, "line": L, "brick": [nx, ny, nz], "ms": tg, "GDoF_s": N / tg / 1e6,
                  "frac_hbm": 32 * N / tg / 1e6 / PEAK, "frac_hbm_moved_112B": 112 * N / tg / 1e6 / PEAK})
            emit({"op": "div", "axis": axis, "line": L, "brick": [nx, ny, nz], "ms": td, "GDoF_s": N / td / 1e6,
                  "frac_hbm": 32 * N / td / 1e6 / PEAK, "frac_hbm_moved_112B": 112 * N / td / 1e6 / PEAK})
            h.close()
            del f, o, v, g3
            torch.cuda.empty_cache()
    # batched general-coefficient tridiagonal solves, 2^24 points per batch
    for L in (64, 128, 256, 512, 1024, 2048):
        nl = (1 << 24) // L
        for layout in ("elem_major", "line_major"):
            shape = (L, nl) if layout == "elem_major" else (nl, L)
            es, ls = (nl, 1) if layout == "elem_major" else (1, L)
            A, B, C, D = (torch.rand(shape, dtype=torch.float64, device="cuda") for _ in range(4))
            B += 2.5
            ptr = [ctypes.c_void_p(t.data_ptr()) for t in (A, B, C, D)]
            tp = tm(lambda: check(LIB.pbx_tdma_periodic_batch_device(L, nl, es, ls, *ptr, None)))
            tf = tm(lambda: check(LIB.pbx_fwd_sweep_batch_device(L, nl, es, ls, *ptr, None)), reps=1, warm=0)
            B.uniform_(2.5, 3.5)
            tb = tm(lambda: check(LIB.pbx_bwd_sweep_batch_device(L, nl, es, ls, *ptr[1:], None)), reps=1, warm=0)
            pts = L * nl
            emit({"op": "tdma_periodic", "line": L, "lines": nl, "layout": layout, "ms": tp, "Gpt_s": pts / tp / 1e6,
                  "frac_hbm": 48 * pts / tp / 1e6 / PEAK})
            emit({"op": "tdma(fwd+bwd)", "line": L, "lines": nl, "layout": layout, "ms": tf + tb,
                  "Gpt_s": pts / (tf + tb) / 1e6, "frac_hbm": 48 * pts / (tf + tb) / 1e6 / PEAK})
            del A, B, C, D
    if a.out:
        with open(a.out, "w") as fh:
            for rec in lines:
                fh.write(json.dumps(rec) + "\n")
    print("\nop             axis line   brick/lines          ms     G(DoF|pt)/s  frac of HBM peak (algorithmic bytes)")
    for r in lines:
        thr = r.get("GDoF_s", r.get("Gpt_s"))
        where = str(r.get("brick", f"{r.get('lines')} {r.get('layout', '')}"))
        print(f"{r['op']:14s} {r.get('axis', '-'):4s} {r['line']:5d}  {where:22s} {r['ms']:8.3f}  {thr:8.2f}  {r.get('frac_hbm', 0):.3f}")
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=27)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-ops", action="store_true")
    a = ap.parse_args()
    run(a.logn, a.skip_ops, a.out)


if __name__ == "__main__":
    main()
