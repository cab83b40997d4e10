"""Generates tests/golden/*.npz: seeded inputs and the CPU oracle's outputs for them.

The reference (Fortran + PETSc) cannot be compiled or imported in this image, and its own tests hold
no stored vectors, so these fixtures are produced by the oracle (oracle/pbx_oracle.c), which is
itself pinned against the reference's analytic known-answer tests (tests/test_oracle_kat.py).
They freeze the oracle's bits: a later change of compiler, flags or oracle source that alters any
result is caught by tests/test_golden.py, and the GPU REFERENCE schedule must reproduce them bit
for bit (tests/test_parity_gpu.py).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as orc  # noqa: E402


def tdma_system(n, rng, periodic):
    a, b, c, x = (rng.random(n) for _ in range(4))
    if not periodic:
        a[0] = 0.0
        c[n - 1] = 0.0
    for i in range(n):
        while abs(b[i]) < abs(a[i]) + abs(c[i]):
            b[i] = 10 * b[i]
    d = b * x + a * np.roll(x, 1) + c * np.roll(x, -1)
    return a, b, c, x, d


def main():
    rng = np.random.default_rng(1234)
    out = {}
    # --- tridsol, n = 33 and 128 ---
    for n in (33, 128):
        for per in (0, 1):
            a, b, c, x, d = tdma_system(n, rng, bool(per))
            key = f"tri_n{n}_p{per}"
            out[key + "_abcd"] = np.stack([a, b, c, d])
            bp, dp = orc.fwd_sweep(a, b, c, d)
            out[key + "_fwd"] = np.stack([bp, dp])
            out[key + "_bwd"] = orc.bwd_sweep(b, c, d)
            out[key + "_tdma"] = orc.tdma(a, b, c, d)
            out[key + "_tdmap"] = orc.tdma_periodic(a, b, c, d)
    # --- 1-D compact operators, n = 37 (odd, not a multiple of anything) and 128 ---
    for n in (37, 128):
        f = rng.uniform(-1, 1, n)
        dx = 1.0 / n
        out[f"l1_n{n}_f"] = f
        out[f"l1_n{n}_grad"] = orc.grad_1d(f, dx)
        out[f"l1_n{n}_div"] = orc.div_1d(f, dx)
        out[f"l1_n{n}_interp"] = orc.interp_1d(f)
        out[f"l1_n{n}_interpdiv"] = orc.interp_1d_div(f)
    # --- 3-D operators on a small non-cubic brick (fast path eligible: multiples of 16) and a
    #     ragged one (reference schedule only) ---
    for tag, shape, dx in (("a", (32, 16, 48), (1 / 32, 0.5 / 16, 2.0 / 48)),
                           ("b", (12, 9, 7), (0.1, 0.2, 0.3))):
        f = np.asfortranarray(rng.uniform(-1, 1, shape))
        v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
        out[f"f3{tag}_f"] = f
        out[f"f3{tag}_v"] = v
        out[f"f3{tag}_dx"] = np.array(dx)
        out[f"f3{tag}_lapl"] = orc.lapl(f, dx)
        out[f"f3{tag}_grad"] = orc.grad(f, dx)
        out[f"f3{tag}_div"] = orc.div(v, dx)
        out[f"f3{tag}_interp"] = orc.interp(f)
        out[f"f3{tag}_interpdiv"] = orc.interp_div(f)
    # --- CG on 16^3, demo recipe (src/example.f90:70-72) ---
    n = 16
    xt = np.asfortranarray(rng.uniform(-1, 1, (n, n, n)))
    dx = (1.0 / n,) * 3
    b = orc.lapl(xt, dx)
    x, its, rnorm, reason, hist = orc.cg_solve(b, dx, rtol=1e-8)
    out["cg16_b"] = b
    out["cg16_x"] = x
    out["cg16_hist"] = hist
    out["cg16_meta"] = np.array([its, reason, rnorm])
    # --- the 2nd-order star (what mfmult applies today) on the two bricks above, and CG on it ---
    for tag in "ab":
        out[f"f3{tag}_star"] = orc.star(out[f"f3{tag}_f"], out[f"f3{tag}_dx"])
    bs = orc.star(xt, dx)
    x, its, rnorm, reason, hist = orc.cg_solve(bs, dx, rtol=1e-8, op=1)
    out["cgstar16_b"] = bs
    out["cgstar16_x"] = x
    out["cgstar16_hist"] = hist
    out["cgstar16_meta"] = np.array([its, reason, rnorm])
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
