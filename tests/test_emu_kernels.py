"""CPU: the library's own kernel sources on the fiber model of tests/emu (TEST INFRASTRUCTURE, see
tests/emu/include/pbx_emu.h), checked against the oracle.  These tests pin KERNEL LOGIC -- tile
geometry, chunk look-back, barriers, TMA box coordinates, slab boundaries -- in a container
without a GPU; the parity claims themselves are made by the `-m gpu` tests on the real kernels
(tests/test_parity_gpu.py), which these mirror at sizes the model runs in seconds."""
import ctypes
import os

import numpy as np
import pytest

import emu_lib
import oracle_lib as orc
from test_parity_gpu import assert_fast_close, tdma_init

EPS = np.finfo(np.float64).eps


def handle(shape, dx, no_tma="0", **kw):
    os.environ["PBX_NO_TMA"] = no_tma
    try:
        return emu_lib.EmuHandle(*shape, dx, **kw)
    finally:
        os.environ.pop("PBX_NO_TMA", None)


def field(shape, seed=1234, ncomp=None):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.uniform(-1, 1, shape if ncomp is None else shape + (ncomp,)))


@pytest.mark.parametrize("shape", [(32, 16, 48), (64, 32, 16), (16, 48, 32), (128, 16, 16), (16, 16, 128)])
@pytest.mark.parametrize("no_tma", ["0", "1"])
def test_emu_lapl_fast_and_reference(shape, no_tma):
    dx = tuple(1.0 / n for n in shape)
    f = field(shape)
    ref = orc.lapl(f, dx)
    lib = emu_lib.load()
    h = handle(shape, dx, no_tma)
    maps0 = lib.pbx_emu_tensor_maps_total()
    out = h.lapl(f)
    if no_tma == "1":
        assert lib.pbx_emu_tensor_maps_total() == maps0   # generic kernels only
    elif shape[1] in (16, 32) and shape[2] in (16, 32, 128):
        assert lib.pbx_emu_tensor_maps_total() - maps0 == 7   # all three TMA kernels ran
    assert_fast_close(out, ref)
    h.set_mode(1)
    assert np.array_equal(h.lapl(f), ref)
    h.close()


@pytest.mark.parametrize("shape", [(64, 32, 64), (16, 1024, 16), (16, 16, 656), (1024, 16, 16)])
def test_emu_tma_and_generic_bit_identical(shape, monkeypatch):
    """includes y / z lines of more than 512 points, which run as overlapping segments, and x lines
    of 1024 points (several warps per line in the TMA x kernel)"""
    monkeypatch.setenv("PBX_TMA_SEG", "1")   # segmented TMA tiles (the default since the proxy-fence fix of round 2)
    dx = tuple(1.0 / n for n in shape)
    f = field(shape, 7)
    outs = []
    for no_tma in ("0", "1"):
        h = handle(shape, dx, no_tma)
        outs.append(h.lapl_dot(f))
        h.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]
    ref = np.vdot(f, outs[0][0])
    assert abs(outs[0][1] - ref) <= 1e-13 * abs(ref)
    orc.set_threads(8)
    try:
        assert_fast_close(outs[0][0], orc.lapl(f, dx))
    finally:
        orc.set_threads(1)


@pytest.mark.parametrize("shape", [(32, 16, 48), (64, 64, 16), (16, 1024, 16), (16, 16, 656)])
def test_emu_grad_div_interp(shape):
    dx = tuple(0.7 / n for n in shape)
    f, v = field(shape, 4321), field(shape, 4322, 3)
    want = (orc.grad(f, dx), orc.div(v, dx), orc.interp(f), orc.interp_div(f))
    h = handle(shape, dx)
    h.set_mode(1)
    got = (h.grad(f), h.div(v), h.interp(f), h.interp(f, +1))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    h.set_mode(0)
    got = (h.grad(f), h.div(v), h.interp(f), h.interp(f, +1))
    for g, w in zip(got, want):
        assert np.max(np.abs(g - w)) <= 1e-13 * np.max(np.abs(w))
    h.close()


@pytest.mark.parametrize("n", [2, 3, 33, 128])
def test_emu_tridsol_bit_exact(n):
    lib = emu_lib.load()
    rng = np.random.default_rng(n)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for per in (False, True):
        a, b, c, x, d = tdma_init(n, rng, per)
        bb, dd = b.copy(), d.copy()
        emu_lib.check(lib, lib.pbx_tdma_host(n, dp(a), dp(bb), dp(c), dp(dd)))
        bo, _ = orc.fwd_sweep(a, b, c, d)
        assert np.array_equal(dd, orc.tdma(a, b, c, d)) and np.array_equal(bb, bo)
        dd = d.copy()
        emu_lib.check(lib, lib.pbx_tdma_periodic_host(n, dp(a), dp(b), dp(c), dp(dd)))
        assert np.array_equal(dd, orc.tdma_periodic(a, b, c, d))


@pytest.mark.parametrize("n", [5, 64, 203])
def test_emu_tridsol_batch_layouts(n):
    """many lines per launch, line-major and element-major layouts, all four routines"""
    lib = emu_lib.load()
    rng = np.random.default_rng(3 + n)
    nl = 37
    for per in (False, True):
        sys_ = [tdma_init(n, rng, per) for _ in range(nl)]
        A, B, C, D = (np.stack([s_[i] for s_ in sys_]) for i in (0, 1, 2, 4))   # [line][i]
        want_t = np.stack([orc.tdma(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_])
        want_p = np.stack([orc.tdma_periodic(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_])
        want_b = np.stack([orc.fwd_sweep(s_[0], s_[1], s_[2], s_[4])[0] for s_ in sys_])
        for layout in ("line_major", "elem_major"):
            tr = (lambda v: v.copy()) if layout == "line_major" else (lambda v: np.ascontiguousarray(v.T))
            es, ls = (1, n) if layout == "line_major" else (nl, 1)
            back = (lambda v: v) if layout == "line_major" else (lambda v: v.T)
            a, b, c, d = (tr(v) for v in (A, B, C, D))
            emu_lib.check(lib, lib.pbx_tdma_batch_device(n, nl, es, ls, emu_lib.ptr(a), emu_lib.ptr(b),
                                                         emu_lib.ptr(c), emu_lib.ptr(d), None))
            assert np.array_equal(back(d), want_t) and np.array_equal(back(b), want_b)
            a, b, c, d = (tr(v) for v in (A, B, C, D))
            emu_lib.check(lib, lib.pbx_tdma_periodic_batch_device(n, nl, es, ls, emu_lib.ptr(a), emu_lib.ptr(b),
                                                                  emu_lib.ptr(c), emu_lib.ptr(d), None))
            assert np.array_equal(back(d), want_p) and np.array_equal(back(b), B)


def test_emu_cg_matches_oracle():
    n = 16
    dx = (2 * np.pi / n,) * 3
    c = (np.arange(n) + 0.5) * dx[0]
    u = np.exp(np.sin(c)[:, None, None] + np.sin(c)[None, :, None] + np.sin(c)[None, None, :])
    b = orc.lapl(np.asfortranarray(u), dx)
    xo, its_o, rn_o, why_o, hist_o = orc.cg_solve(b, dx, rtol=1e-8)
    h = handle((n, n, n), dx)
    x, its, rn, why, hist = h.cg_solve(b, rtol=1e-8)
    h.close()
    assert why == why_o == 2 and abs(its - its_o) <= 1
    m = min(len(hist), len(hist_o))
    assert np.allclose(hist[:m], hist_o[:m], rtol=1e-6)
    assert np.max(np.abs(x - xo)) <= 1e-6 * np.max(np.abs(xo))


@pytest.mark.parametrize("shape,P", [((32, 16, 128), 2), ((16, 16, 256), 4), ((16, 528, 128), 2), ((16, 16, 192), 2)])
@pytest.mark.parametrize("no_tma", ["0", "1"])
def test_emu_slabs_match_single_brick(shape, P, no_tma):
    nx, ny, nz = shape
    nzl = nz // P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    f = field(shape, 5)
    whole = handle(shape, dx, no_tma)
    ref = whole.lapl(f)
    slabs = [handle((nx, ny, nzl), dx, no_tma, slab=(r, P)) for r in range(P)]
    for r, h in enumerate(slabs):
        h.slab_phase1(np.asfortranarray(f[:, :, r * nzl:(r + 1) * nzl]))
    emu_lib.EmuHandle.slab_exchange_local(slabs)
    out = np.concatenate([h.slab_phase2() for h in slabs], axis=2)
    assert np.max(np.abs(out - ref)) <= 1e-13 * np.max(np.abs(ref))
    for h in slabs + [whole]:
        h.close()


@pytest.mark.parametrize("shape,P", [((32, 16, 128), 2), ((16, 32, 192), 3)])
def test_emu_slab_grad_div_interp(shape, P):
    """grad / div / interp on P slabs (phase 1, exchange, phase 2) against the whole brick"""
    from poissbox_b200 import _lib

    nx, ny, nz = shape
    nzl = nz // P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    f, v = field(shape, 11), field(shape, 12, 3)
    orc.set_threads(8)
    try:
        want = {_lib.OP_GRAD: orc.grad(f, dx), _lib.OP_DIV: orc.div(v, dx),
                _lib.OP_INTERP: orc.interp(f), _lib.OP_INTERP_DIV: orc.interp_div(f)}
    finally:
        orc.set_threads(1)
    slabs = [handle((nx, ny, nzl), dx, slab=(r, P)) for r in range(P)]
    for op, w in want.items():
        src = v if op == _lib.OP_DIV else f
        for r, h in enumerate(slabs):
            h.slab_op_phase1(op, np.asfortranarray(src[:, :, r * nzl:(r + 1) * nzl]))
        emu_lib.EmuHandle.slab_exchange_local(slabs)
        out = np.concatenate([h.slab_op_phase2(op) for h in slabs], axis=2)
        assert np.max(np.abs(out - w)) <= 1e-13 * np.max(np.abs(w)), op
    for h in slabs:
        h.close()


@pytest.mark.parametrize("shape", [(12, 9, 7), (32, 16, 48), (3, 3, 3), (40, 33, 70)])
def test_emu_star_bit_exact(shape):
    """the 2nd-order star (what mfmult applies today): bit-identical to the oracle, any extents"""
    dx = (0.1, 0.25, 0.37)
    f = field(shape, 3)
    h = handle(shape, dx)
    assert np.array_equal(h.star(f), orc.star(f, dx))
    h.set_operator(1)
    assert np.array_equal(h.mult(f), orc.star(f, dx))
    h.set_operator(0)
    if all(n % 16 == 0 for n in shape):
        assert np.array_equal(h.mult(f), h.lapl(f))
    h.close()


def test_emu_star_slabs_and_cg():
    from poissbox_b200 import _lib

    shape, P, nzl, dx = (16, 16, 128), 2, 64, (0.1, 0.2, 0.3)
    f = field(shape, 8)
    slabs = [handle((16, 16, nzl), dx, slab=(r, P)) for r in range(P)]
    for r, h in enumerate(slabs):
        h.slab_op_phase1(_lib.OP_STAR, np.asfortranarray(f[:, :, r * nzl:(r + 1) * nzl]))
    emu_lib.EmuHandle.slab_exchange_local(slabs)
    out = np.concatenate([h.slab_op_phase2(_lib.OP_STAR) for h in slabs], axis=2)
    assert np.array_equal(out, orc.star(f, dx))
    for h in slabs:
        h.close()
    # CG on the star operator = the reference's solve() as it stands today
    n = 16
    dx = (2 * np.pi / n,) * 3
    c = (np.arange(n) + 0.5) * dx[0]
    u = np.exp(np.sin(c)[:, None, None] + np.sin(c)[None, :, None] + np.sin(c)[None, None, :])
    b = orc.star(np.asfortranarray(u), dx)
    xo, its_o, _, why_o, _ = orc.cg_solve(b, dx, rtol=1e-8, op=1)
    h = handle((n, n, n), dx)
    h.set_operator(1)
    x, its, _, why, _ = h.cg_solve(b, rtol=1e-8)
    h.close()
    assert why == why_o == 2 and abs(its - its_o) <= 1
    assert np.max(np.abs(x - xo)) <= 1e-8 * np.max(np.abs(xo))


@pytest.mark.parametrize("shape,dx", [((16, 16, 16), (0.1, 0.1, 0.1)), ((32, 16, 48), (0.1, 0.15, 0.07)),
                                      ((24, 16, 20), (0.3, 0.2, 0.25))])
def test_emu_mg_vcycle_matches_model(shape, dx):
    """the multigrid preconditioner (pbx_mg.cu) against its numpy model, and its symmetry and
    positive definiteness on the zero-mean subspace (what CG needs of a preconditioner)"""
    import mg_model as mg
    from poissbox_b200 import _lib

    rng = np.random.default_rng(0)
    r = np.asfortranarray(rng.standard_normal(shape)) + 0.3
    q = np.asfortranarray(rng.standard_normal(shape))
    for nu in (1, 2):
        h = handle(shape, dx)
        h.set_pc(_lib.PC_MG, nu)
        z = h.pc_apply(r)
        zm = mg.pc_apply(r, dx, nu)
        assert np.max(np.abs(z - zm)) <= 1e-13 * np.max(np.abs(zm))
        assert abs(z.mean()) <= 1e-13 * np.max(np.abs(z))
        zq = h.pc_apply(q)
        a, b = np.vdot(q - q.mean(), z), np.vdot(zq, r - r.mean())
        assert abs(a - b) <= 1e-12 * max(abs(a), abs(b))
        assert np.vdot(r - r.mean(), z) > 0
        h.close()
    h = handle(shape, dx)   # PC none: z = r - mean(r)
    assert np.max(np.abs(h.pc_apply(r) - (r - r.mean()))) <= 1e-15 * np.max(np.abs(r))
    h.close()


@pytest.mark.parametrize("op", [0, 1])
def test_emu_pcg_matches_model(op):
    """KSPCG with the multigrid preconditioner: iteration count and history of the numpy model"""
    import mg_model as mg
    from poissbox_b200 import _lib

    n = 16
    hh = 2 * np.pi / n
    dx = (hh,) * 3
    c = (np.arange(n) + 0.5) * hh
    u = np.exp(np.sin(c)[:, None, None] + np.sin(c)[None, :, None] + np.sin(c)[None, None, :])
    apply_a = (lambda v: orc.lapl(np.asfortranarray(v), dx)) if op == 0 else (lambda v: orc.star(np.asfortranarray(v), dx))
    b = apply_a(u)
    xm, itm, histm = mg.pcg(apply_a, b, lambda r: mg.pc_apply(r, dx, 2))
    h = handle((n, n, n), dx)
    h.set_operator(op)
    h.set_pc(_lib.PC_MG, 2)
    x, its, rn, why, hist = h.cg_solve(b, rtol=1e-8)
    assert why == 2 and its == itm
    assert np.allclose(hist, histm[: its + 1], rtol=1e-9)
    du = u - u.mean()
    assert np.linalg.norm((x - x.mean()) - du) <= 1e-6 * np.linalg.norm(du)
    h.set_pc(_lib.PC_NONE)
    _, its0, _, why0, _ = h.cg_solve(b, rtol=1e-8)
    assert why0 == 2 and its0 > 3 * its   # the unpreconditioned solve still works and is much longer
    h.close()


def test_emu_no_device_is_an_error():
    """the harness honours the product's rule: no device, no result (PBX_ERR_CUDA)"""
    lib = emu_lib.load()
    os.environ["PBX_EMU_NO_DEVICE"] = "1"
    try:
        h = ctypes.c_void_p()
        from poissbox_b200 import _lib
        rc = lib.pbx_create(16, 16, 16, _lib._d3(1.0, 1.0, 1.0), 0, None, ctypes.byref(h))
        assert rc == _lib.PBX_ERR_CUDA
    finally:
        os.environ.pop("PBX_EMU_NO_DEVICE", None)


@pytest.mark.parametrize("n,nl,pad", [(5, 37, 1), (64, 37, 0), (203, 70, 1), (128, 32, 0), (48, 3, 4), (16, 33, 0)])
def test_emu_tridsol_line_major_tma(n, nl, pad, monkeypatch):
    """PBX_TDMA_TMA=1: contiguous lines travel as swizzled TMA tiles (pbx_tdma_tma.cu), a warp per 32
    lines, results written in place into the tiles -- same bits as the oracle, for lines that do not
    fill the last 16-point block, batches that do not fill the last CTA and padded line strides"""
    lib = emu_lib.load()
    monkeypatch.setenv("PBX_TDMA_TMA", "1")
    rng = np.random.default_rng(100 * n + nl)
    ls = n + pad + ((n + pad) & 1)                 # even line stride (16-byte aligned rows)
    for per in (False, True):
        sys_ = [tdma_init(n, rng, per) for _ in range(nl)]
        want_t = np.stack([orc.tdma(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_])
        want_p = np.stack([orc.tdma_periodic(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_])
        want_b = np.stack([orc.fwd_sweep(s_[0], s_[1], s_[2], s_[4])[0] for s_ in sys_])

        def padded(i):
            v = emu_lib.new_field((ls, nl))        # Fortran order: [line][ls] in memory, poison in the padding
            v[:n, :] = np.stack([s_[i] for s_ in sys_]).T
            return v

        a, b, c, d = (padded(i) for i in (0, 1, 2, 4))
        maps0 = lib.pbx_emu_tensor_maps_total()
        emu_lib.check(lib, lib.pbx_tdma_batch_device(n, nl, 1, ls, emu_lib.ptr(a), emu_lib.ptr(b), emu_lib.ptr(c),
                                                     emu_lib.ptr(d), None))
        # odd n stays on the generic kernels: TMA moves 16-byte units and would touch the element behind the line
        assert lib.pbx_emu_tensor_maps_total() - maps0 == (7 if n % 2 == 0 else 0), "the TMA kernels did not run"
        assert np.array_equal(d[:n].T, want_t) and np.array_equal(b[:n].T, want_b)
        assert np.all(d[n:] == 73.29) and np.all(b[n:] == 73.29), "padding between the lines touched"
        a, b, c, d = (padded(i) for i in (0, 1, 2, 4))
        b0 = b.copy()
        maps0 = lib.pbx_emu_tensor_maps_total()
        emu_lib.check(lib, lib.pbx_tdma_periodic_batch_device(n, nl, 1, ls, emu_lib.ptr(a), emu_lib.ptr(b),
                                                              emu_lib.ptr(c), emu_lib.ptr(d), None))
        if n % 2 == 0:
            assert lib.pbx_emu_tensor_maps_total() - maps0 == 6, "the TMA kernel did not run"
        assert np.array_equal(d[:n].T, want_p) and np.array_equal(b, b0)
        assert np.all(d[n:] == 73.29)


def test_emu_lapl_host_batch():
    """pbx_lapl_host_batch: several fields of one box per call (double-buffered staging slots, three
    streams on the GPU): every output equals the single-field call's, for more fields than slots"""
    from poissbox_b200 import _lib

    lib = emu_lib.load()
    n, dx = (32, 16, 16), (0.1, 0.2, 0.3)
    rng = np.random.default_rng(0)
    fs = [emu_lib.aligned(np.asfortranarray(rng.uniform(-1, 1, n))) for _ in range(5)]
    outs = [emu_lib.new_field(n) for _ in fs]
    pin = (ctypes.c_void_p * 5)(*[f.ctypes.data for f in fs])
    pout = (ctypes.c_void_p * 5)(*[o.ctypes.data for o in outs])
    for mode in (0, 1):
        emu_lib.check(lib, lib.pbx_lapl_host_batch(*n, 5, pin, _lib._d3(*dx), pout, mode))
        h = handle(n, dx)
        h.set_mode(mode)
        assert all(np.array_equal(h.lapl(f), o) for f, o in zip(fs, outs))
        h.close()
    assert lib.pbx_lapl_host_batch(*n, 0, pin, _lib._d3(*dx), pout, 0) != 0


@pytest.mark.parametrize("shape", [(32, 512, 16), (16, 64, 512), (48, 256, 32), (16, 32, 512), (32, 128, 16)])
def test_emu_yz_rot_bit_identical(shape, monkeypatch):
    """PBX_YZ_ROT=1: swizzled y/z tiles read in a bank-conflict-free order and put back in place by
    register swaps -- the same bits as the unswizzled TMA kernels and the generic ones, dot included"""
    dx = tuple(1.0 / n for n in shape)
    f = field(shape, 21)
    lib = emu_lib.load()
    h = handle(shape, dx)
    monkeypatch.setenv("PBX_YZ_ROT", "0")     # the unswizzled tile reads the (now default) rotated ones replaced
    ref, dref = h.lapl_dot(f)
    monkeypatch.setenv("PBX_YZ_ROT", "1")
    maps0 = lib.pbx_emu_tensor_maps3_swizzled_total()
    out, dot = h.lapl_dot(f)
    # swizzled 3-D maps: two per rotated pass (y always; z when the tile holds one line per row)
    assert lib.pbx_emu_tensor_maps3_swizzled_total() - maps0 == (4 if shape[2] == 512 else 2)
    assert np.array_equal(out, ref) and dot == dref
    monkeypatch.delenv("PBX_YZ_ROT")
    g = handle(shape, dx, no_tma="1")
    assert np.array_equal(g.lapl(f), ref)
    h.close()
    g.close()


@pytest.mark.parametrize("rtol,maxit", [(1e-2, 10000), (1e-8, 3), (1e-8, 1), (0.5, 10000)])
def test_emu_cg_last_step_reaches_x(rtol, maxit):
    """x += a p rides with the p update (k_pupdate_x); the iteration that converges or hits max_it must
    still apply its step although the status word is already set: x equals the oracle CG's to rounding"""
    n = 16
    dx = (2 * np.pi / n,) * 3
    b = orc.lapl(field((n, n, n), 5), dx)
    xo, ito, _, whyo, _ = orc.cg_solve(b, dx, rtol=rtol, maxit=maxit)
    h = handle((n, n, n), dx)
    for _ in range(2):   # a second solve on the same handle starts from a clean iteration stamp
        x, it, _, why, _ = h.cg_solve(b, rtol=rtol, maxit=maxit)
        assert (it, why) == (ito, whyo)
        assert np.max(np.abs(x - xo)) <= 1e-12 * np.max(np.abs(xo))
    h.close()


def test_emu_ksp_options(capfd):
    """pbx_ksp_solve_device: the solve configured by the PETSc option names of the reference's README
    (-ksp_type cg -pc_type ... -ksp_rtol ... -ksp_monitor -ksp_converged_reason)"""
    from poissbox_b200 import _lib

    n = 16
    dx = (2 * np.pi / n,) * 3
    b = orc.lapl(field((n, n, n), 5), dx)
    h = handle((n, n, n), dx)
    x0, it0, rn0, why0, hist0 = h.cg_solve(b, rtol=1e-3)
    x, it, rn, why = h.ksp_solve(b, "-ksp_type cg -pc_type none -ksp_rtol 1e-3 -ksp_monitor -ksp_converged_reason")
    out = capfd.readouterr().out
    assert (it, rn, why) == (it0, rn0, why0) and np.array_equal(x, x0)
    lines = [ln for ln in out.splitlines() if "KSP Residual norm" in ln]
    assert len(lines) == it + 1 and lines[0].split()[0] == "0" and float(lines[-1].split()[-1]) == pytest.approx(rn, rel=1e-11)
    assert f"Linear solve converged due to CONVERGED_RTOL iterations {it}" in out
    # PETSc's defaults when nothing is said; unknown options are ignored
    assert h.ksp_solve(b, "-log_view -da_grid_x 16")[1:] == h.cg_solve(b)[1:4]
    # -pc_type gamg selects the multigrid stand-in and stays selected; -ksp_max_it bounds the solve
    xm, itm, _, whym = h.ksp_solve(b, "-pc_type gamg -ksp_rtol 1e-6 -mg_levels_ksp_max_it 2")
    h2 = handle((n, n, n), dx)
    h2.set_pc(_lib.PC_MG, 2)
    assert (itm, whym) == h2.cg_solve(b, rtol=1e-6)[1:4:2]
    assert h.ksp_solve(b, "-ksp_max_it 2 -ksp_rtol 1e-12")[1::2] == (2, -3)
    for bad, code in (("-ksp_type gmres", 4), ("-pc_type ilu", 4), ("-ksp_rtol", 1), ("-ksp_rtol abc", 1),
                      ("-ksp_max_it -3", 1)):
        with pytest.raises(_lib.PbxError) as e:
            h.ksp_solve(b, bad)
        assert e.value.code == code
    h.close()
    h2.close()


@pytest.mark.parametrize("shape", [(64, 32, 64), (32, 512, 16), (16, 16, 512), (512, 16, 16), (48, 64, 32), (16, 256, 16),
                                   (1024, 16, 16), (2048, 16, 16), (16, 1024, 16), (16, 16, 2048), (16, 576, 16)])
def test_emu_lineop_tma_bit_identical(shape, monkeypatch):
    """PBX_LINEOP_TMA=1: grad / div / interp through the TMA-pipelined line-operator kernels (two tile
    stages, x direction by shuffles) -- the same bits as the generic line-operator kernels"""
    dx = tuple(0.7 / n for n in shape)
    f = field(shape, 31)
    vec = np.asfortranarray(np.random.default_rng(32).uniform(-1, 1, shape + (3,)))
    lib = emu_lib.load()
    h = handle(shape, dx)
    # segmented y / z lines: the TMA kernels are off by default (open defect on the B200, pbx_fast_tma.cu); the
    # harness keeps exercising their logic
    monkeypatch.setenv("PBX_TMA_SEG", "1")
    monkeypatch.setenv("PBX_LINEOP_TMA", "0")  # the generic line-operator kernels
    want = [h.grad(f), h.div(vec), h.interp(f), h.interp(f, +1)]
    monkeypatch.setenv("PBX_LINEOP_TMA", "1")
    maps0 = lib.pbx_emu_tensor_maps_total()
    lib.pbx_launch_count.restype = ctypes.c_longlong
    l0 = lib.pbx_launch_count(h._h)
    got = [h.grad(f), h.div(vec), h.interp(f), h.interp(f, +1)]
    assert lib.pbx_emu_tensor_maps_total() > maps0, "the TMA line operators did not run"
    # grad: two operator pairs share a launch; div: its two sums are formed inside a two-input launch
    assert lib.pbx_launch_count(h._h) - l0 == 6 + 6 + 3 + 3
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    h.close()


@pytest.mark.parametrize("shape", [(16, 48, 192), (32, 384, 16), (16, 320, 48), (16, 96, 80), (32, 16, 272),
                                   (48, 16, 16), (192, 32, 16), (1600, 16, 16), (96, 48, 80)])
def test_emu_tma_any_chunk_count(shape, monkeypatch):
    """PBX_TMA_ANY_T=1: lines whose chunk count is not a power of two (x: any multiple of 16 up to 4096;
    y / z: 48 ... 384 points) on the TMA kernels -- a tile holds the whole lines that fit, the remaining
    threads idle; the same bits as the generic kernels, Laplacian with its fused dot and the line
    operators"""
    dx = tuple(0.9 / n for n in shape)
    f = field(shape, 41)
    lib = emu_lib.load()
    g = handle(shape, dx, no_tma="1")
    ref, dref = g.lapl_dot(f)
    gref = g.grad(f)
    monkeypatch.setenv("PBX_TMA_ANY_T", "1")
    h = handle(shape, dx)
    maps0 = lib.pbx_emu_tensor_maps_total()
    out, dot = h.lapl_dot(f)
    nmaps = lib.pbx_emu_tensor_maps_total() - maps0
    odd = lambda n: (n // 16) & (n // 16 - 1) != 0
    assert nmaps >= 3 * odd(shape[0]) + 2 * odd(shape[1]) + 2 * odd(shape[2]), "the TMA kernels did not run"
    # the field carries the same bits; the fused dot is summed over 256 threads (idle ones adding zeros)
    # instead of the generic kernel's 8 T G, i.e. in another association
    assert np.array_equal(out, ref) and abs(dot - dref) <= 1e-13 * abs(dref)
    monkeypatch.setenv("PBX_LINEOP_TMA", "1")
    assert np.array_equal(h.grad(f), gref)
    h.close()
    g.close()


def test_emu_fused_reduction_tails(monkeypatch):
    """PBX_FUSE_TAIL=1: the kernels that write per-CTA partial sums (z pass with its fused dot, residual
    update) reduce them in their last CTA and run the CG's scalar step there: five launches per
    iteration instead of nine, the same iterations and the same x"""
    n = 16
    dx = (2 * np.pi / n,) * 3
    b = orc.lapl(field((n, n, n), 5), dx)
    h = handle((n, n, n), dx)
    lib = h.lib
    lib.pbx_launch_count.restype = ctypes.c_longlong
    x0, it0, rn0, why0, hist0 = h.cg_solve(b, rtol=1e-6)
    out0, dot0 = h.lapl_dot(b)
    monkeypatch.setenv("PBX_FUSE_TAIL", "1")
    l0 = lib.pbx_launch_count(h._h)
    x1, it1, rn1, why1, hist1 = h.cg_solve(b, rtol=1e-6)
    per_it = (lib.pbx_launch_count(h._h) - l0) / it1
    assert (it1, why1) == (it0, why0) and per_it < 5.5
    assert np.allclose(hist1, hist0, rtol=1e-12) and np.max(np.abs(x1 - x0)) <= 1e-12 * np.max(np.abs(x0))
    for kw in (dict(rtol=0.5), dict(rtol=1e-8, maxit=3)):      # early exits: the last step still reaches x
        monkeypatch.delenv("PBX_FUSE_TAIL")
        a = h.cg_solve(b, **kw)
        monkeypatch.setenv("PBX_FUSE_TAIL", "1")
        c = h.cg_solve(b, **kw)
        assert a[1:4:2] == c[1:4:2] and np.max(np.abs(a[0] - c[0])) <= 1e-12 * np.max(np.abs(a[0]))
    out1, dot1 = h.lapl_dot(b)
    assert np.array_equal(out1, out0) and abs(dot1 - dot0) <= 1e-13 * abs(dot0)
    h.close()



def test_emu_fused_tail_on_rotated_z_kernel(monkeypatch):
    """512-point z lines run on the rotated-tile z kernel; with the reduction tail fused into it (default) the CG takes
    the same iterations to the same x as with the reductions in launches of their own, in five launches per iteration
    instead of nine"""
    shape = (16, 16, 512)
    dx = tuple(2 * np.pi / s for s in shape)
    b = orc.lapl(field(shape, 8), dx)
    h = handle(shape, dx)
    lib = h.lib
    lib.pbx_launch_count.restype = ctypes.c_longlong

    def run(maxit):
        l0 = lib.pbx_launch_count(h._h)
        res = h.cg_solve(b, rtol=1e-30, maxit=maxit)
        return res, lib.pbx_launch_count(h._h) - l0

    out = {}
    for ft in ("0", "1"):
        monkeypatch.setenv("PBX_FUSE_TAIL", ft)
        (_, _, _, _, _), n3 = run(3)
        res, n6 = run(6)
        out[ft] = (res, (n6 - n3) / 3)
    (x0, it0, _, why0, hist0), per0 = out["0"]
    (x1, it1, _, why1, hist1), per1 = out["1"]
    assert (it1, why1) == (it0, why0) and (per0, per1) == (9, 5)
    assert np.allclose(hist1, hist0, rtol=1e-12) and np.max(np.abs(x1 - x0)) <= 1e-12 * np.max(np.abs(x0))
    h.close()
