"""GPU parity tests: the CUDA path, called through the C ABI (libpbx.so), against the CPU oracle.

Bars (BASELINE.json north_star; SURVEY 7 "hard parts"):
  * REFERENCE schedule: BIT-IDENTICAL to the oracle (same operations in the same order, no FMA).
  * FAST schedule: max|delta| <= 1e-12 * max|reference| (it re-associates the computation, so it
    cannot be bit-identical; the measured agreement is ~1e-15), and per-element relative
    difference <= 1e-12 wherever |reference| >= 1e-3 * max|reference|.
  * tridsol batches: bit-identical to the oracle.
The known-answer tests of the reference are restated through the host-side mirror of its module
interface (poissbox_b200.compact_schemes / tridsol), so they read like the reference's own tests.
"""
import os

import numpy as np
import pytest

import oracle_lib as orc
import poissbox_b200 as pbx
from poissbox_b200 import compact_schemes as cs
from poissbox_b200 import tridsol

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
FAST_TOL = 1e-12


def assert_fast_close(got, ref):
    scale = np.max(np.abs(ref))
    d = np.abs(got - ref)
    assert np.max(d) <= FAST_TOL * scale, f"max|d|/max|ref| = {np.max(d) / scale:.3e}"
    big = np.abs(ref) >= 1e-3 * scale
    assert np.max(d[big] / np.abs(ref[big])) <= FAST_TOL


# ------------------------------------------------------------------------------------ tridsol
def tdma_init(n, rng, periodic=False):
    a, b, c, x = (rng.random(n) for _ in range(4))
    if not periodic:
        a[0] = 0.0
        c[n - 1] = 0.0
    for i in range(n):
        while abs(b[i]) < abs(a[i]) + abs(c[i]):
            b[i] = 10 * b[i]
    d = b * x + a * np.roll(x, 1) + c * np.roll(x, -1)
    return a, b, c, x, d


@pytest.mark.parametrize("n", [2, 3, 33, 128, 2048])
def test_tridsol_bit_exact(n):
    rng = np.random.default_rng(n)
    for per in (False, True):
        a, b, c, x, d = tdma_init(n, rng, per)
        # fwd_sweep / bwd_sweep / tdma (tests/tridiag/test_tdma_sweeps.f90, test_tdma.f90)
        bo, do = orc.fwd_sweep(a, b, c, d)
        bg, dg = tridsol.fwd_sweep(a.copy(), b.copy(), c.copy(), d.copy())
        assert np.array_equal(bo, bg) and np.array_equal(do, dg)
        assert np.array_equal(orc.bwd_sweep(b, c, d), tridsol.bwd_sweep(b.copy(), c.copy(), d.copy()))
        bb = b.copy()
        got = tridsol.tdma(a.copy(), bb, c.copy(), d.copy())
        assert np.array_equal(orc.tdma(a, b, c, d), got)
        assert np.array_equal(bb, bo)   # b overwritten with the pivots (tridsol.f90:92)
        # tdma_periodic (tests/tridiag/test_tdma_periodic.f90)
        bb = b.copy()
        got = tridsol.tdma_periodic(a.copy(), bb, c.copy(), d.copy())
        assert np.array_equal(orc.tdma_periodic(a, b, c, d), got)
        assert np.array_equal(bb, b)
        if n >= 3:
            err = np.sqrt(np.mean((got - x) ** 2)) if per else 0.0
            assert err <= 4 * EPS * np.sqrt(np.mean(x**2)) * max(1, n / 128)


def test_tridsol_golden():
    for n in (33, 128):
        for per in (0, 1):
            k = f"tri_n{n}_p{per}"
            a, b, c, d = (np.ascontiguousarray(v) for v in G[k + "_abcd"])
            assert np.array_equal(tridsol.tdma(a.copy(), b.copy(), c.copy(), d.copy()), G[k + "_tdma"])
            assert np.array_equal(tridsol.tdma_periodic(a.copy(), b.copy(), c.copy(), d.copy()), G[k + "_tdmap"])


def test_tridsol_batch_strided():
    """many lines in one launch, both memory layouts (line-major and element-major)"""
    import ctypes

    import torch

    rng = np.random.default_rng(3)
    n, nl = 64, 300
    sys_ = [tdma_init(n, rng, True) for _ in range(nl)]
    A, B, C, D = (np.stack([s[i] for s in sys_]) for i in (0, 1, 2, 4))   # [line][i]
    want = np.stack([orc.tdma_periodic(*s[:3], s[4]) for s in sys_])
    for layout in ("line_major", "elem_major"):
        if layout == "line_major":
            arrs = [torch.from_numpy(v.copy()).cuda() for v in (A, B, C, D)]
            es, ls = 1, n
        else:
            arrs = [torch.from_numpy(np.ascontiguousarray(v.T)).cuda() for v in (A, B, C, D)]
            es, ls = nl, 1
        ptr = [ctypes.c_void_p(t.data_ptr()) for t in arrs]
        pbx.check(pbx.LIB.pbx_tdma_periodic_batch_device(n, nl, es, ls, *ptr, None))
        torch.cuda.synchronize()
        got = arrs[3].cpu().numpy()
        got = got if layout == "line_major" else got.T
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,nl,pad", [(5, 37, 1), (6, 37, 2), (64, 300, 0), (203, 70, 1), (204, 70, 0), (512, 4096, 0),
                                      (2048, 33, 2)])
def test_tridsol_line_major_tma(n, nl, pad, monkeypatch):
    """PBX_TDMA_TMA=1: contiguous lines as swizzled TMA tiles, same bits as the generic kernels and the oracle"""
    import ctypes

    import torch

    rng = np.random.default_rng(100 * n + nl)
    ls = n + pad + ((n + pad) & 1)
    for per in (False, True):
        sys_ = [tdma_init(n, rng, per) for _ in range(min(nl, 64))]
        reps = (nl + len(sys_) - 1) // len(sys_)
        host = []
        for i in (0, 1, 2, 4):
            v = np.full((nl, ls), 73.29)
            v[:, :n] = np.tile(np.stack([s_[i] for s_ in sys_]), (reps, 1))[:nl]
            host.append(v)
        want_t = np.tile(np.stack([orc.tdma(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_]), (reps, 1))[:nl]
        want_p = np.tile(np.stack([orc.tdma_periodic(s_[0], s_[1], s_[2], s_[4]) for s_ in sys_]), (reps, 1))[:nl]
        for fn, want in ((pbx.LIB.pbx_tdma_batch_device, want_t), (pbx.LIB.pbx_tdma_periodic_batch_device, want_p)):
            got = {}
            for tma in ("0", "1"):
                monkeypatch.setenv("PBX_TDMA_TMA", tma)   # "0": the generic thread-per-line kernels
                arrs = [torch.from_numpy(v.copy()).cuda() for v in host]
                pbx.check(fn(n, nl, 1, ls, *[ctypes.c_void_p(t.data_ptr()) for t in arrs], None))
                torch.cuda.synchronize()
                got[tma] = (arrs[1].cpu().numpy(), arrs[3].cpu().numpy())
            assert np.array_equal(got["1"][1][:, :n], want)
            assert np.array_equal(got["0"][0], got["1"][0]) and np.array_equal(got["0"][1], got["1"][1])


def test_lapl_host_batch():
    rng = np.random.default_rng(8)
    n, dx = (64, 32, 48), (0.1, 0.2, 0.3)
    fs = [np.asfortranarray(rng.uniform(-1, 1, n)) for _ in range(5)]
    for mode in (pbx.MODE_FAST, pbx.MODE_REFERENCE):
        outs = cs.lapl_batch(fs, dx, mode=mode)
        for f, o in zip(fs, outs):
            assert np.array_equal(o, cs.lapl(f, dx, mode=mode))


@pytest.mark.parametrize("shape", [(512, 512, 512), (64, 256, 512), (32, 128, 64)])
def test_yz_rot_bit_identical(shape, monkeypatch):
    import torch

    nx, ny, nz = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    h = pbx.Handle(nx, ny, nz, (1.0 / nx, 1.0 / ny, 1.0 / nz))
    monkeypatch.setenv("PBX_YZ_ROT", "0")     # the unswizzled tile reads the default replaced
    ref, dref = h.lapl_dot(f)
    monkeypatch.setenv("PBX_YZ_ROT", "1")
    out, dot = h.lapl_dot(f)
    torch.cuda.synchronize()
    assert torch.equal(out, ref) and dot.item() == dref.item()
    h.close()


@pytest.mark.parametrize("shape", [(256, 256, 256), (64, 512, 32), (48, 64, 512)])
def test_lineop_tma_bit_identical(shape, monkeypatch):
    import torch

    nx, ny, nz = shape
    g = torch.Generator(device="cuda").manual_seed(4)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v = torch.rand((3, nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    h = pbx.Handle(nx, ny, nz, (0.7 / nx, 0.7 / ny, 0.7 / nz))
    monkeypatch.setenv("PBX_LINEOP_TMA", "0")  # the generic line-operator kernels the default replaced
    want = [h.grad(f), h.div(v), h.interp(f), h.interp(f, +1)]
    monkeypatch.setenv("PBX_LINEOP_TMA", "1")
    got = [h.grad(f), h.div(v), h.interp(f), h.interp(f, +1)]
    torch.cuda.synchronize()
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    h.close()


@pytest.mark.parametrize("shape", [(384, 384, 384), (48, 320, 192), (1600, 96, 48)])
def test_tma_any_chunk_count(shape, monkeypatch):
    import torch

    nx, ny, nz = shape
    g = torch.Generator(device="cuda").manual_seed(6)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    h = pbx.Handle(nx, ny, nz, (0.9 / nx, 0.9 / ny, 0.9 / nz))
    monkeypatch.setenv("PBX_TMA_ANY_T", "0")
    ref, dref = h.lapl_dot(f)           # generic kernels for these extents
    monkeypatch.setenv("PBX_TMA_ANY_T", "1")
    out, dot = h.lapl_dot(f)
    torch.cuda.synchronize()
    assert torch.equal(out, ref) and abs(dot.item() - dref.item()) <= 1e-13 * abs(dref.item())
    h.close()


# ------------------------------------------------------------------------------------ 1-D operators
@pytest.mark.parametrize("n", [3, 4, 5, 37, 128, 1000])
def test_lines_bit_exact(n):
    rng = np.random.default_rng(n)
    f = rng.uniform(-1, 1, n)
    dx = 0.37 / n
    assert np.array_equal(cs.grad_1d(f, dx), orc.grad_1d(f, dx))
    assert np.array_equal(cs.div_1d(f, dx), orc.div_1d(f, dx))
    assert np.array_equal(cs.interp_1d(f), orc.interp_1d(f))
    assert np.array_equal(cs.interp_1d_div(f), orc.interp_1d_div(f))


def test_lines_golden():
    for n in (37, 128):
        f = G[f"l1_n{n}_f"]
        assert np.array_equal(cs.grad_1d(f, 1.0 / n), G[f"l1_n{n}_grad"])
        assert np.array_equal(cs.div_1d(f, 1.0 / n), G[f"l1_n{n}_div"])
        assert np.array_equal(cs.interp_1d(f), G[f"l1_n{n}_interp"])
        assert np.array_equal(cs.interp_1d_div(f), G[f"l1_n{n}_interpdiv"])


def test_grad_1d_kat():
    """tests/grad/test_grad_1d.f90:53-134 and tests/div/test_div_1d.f90:53-134 through the mirror"""
    n = 128
    dx = 2 * np.pi / n
    f = np.full(n, 2.8170923)
    df = np.full(n, 73.29)
    cs.grad_1d(f, dx, df)
    assert np.sqrt(np.sum(df**2) / n) <= 100 * EPS
    assert np.sqrt(np.sum((cs.interp_1d(f) - f) ** 2) / n) <= 100 * EPS
    xc, xv = (np.arange(n) + 0.5) * dx, np.arange(n) * dx
    assert np.sqrt(np.mean((cs.grad_1d(np.sin(xc), dx) - np.cos(xv)) ** 2)) <= 1e-11
    assert np.sqrt(np.mean((cs.interp_1d(np.sin(xc)) - np.sin(xv)) ** 2)) <= 1e-11
    assert np.sqrt(np.mean((cs.div_1d(np.sin(xv), dx) - np.cos(xc)) ** 2)) <= 1e-11
    assert np.sqrt(np.mean((cs.interp_1d_div(np.sin(xv)) - np.sin(xc)) ** 2)) <= 1e-11


def test_size_mismatch_stop_7():
    with pytest.raises(pbx.SizeMismatch):
        cs.grad_1d(np.zeros(8), 0.1, df=np.zeros(7))


# ------------------------------------------------------------------------------------ 3-D operators
SHAPES = [((32, 16, 48), (1 / 32, 0.5 / 16, 2.0 / 48)), ((12, 9, 7), (0.1, 0.2, 0.3)),
          ((64, 64, 64), (1 / 64,) * 3), ((3, 3, 3), (1.0, 1.0, 1.0))]


@pytest.mark.parametrize("shape,dx", SHAPES)
def test_fields_reference_bit_exact(shape, dx):
    rng = np.random.default_rng(1234)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
    assert np.array_equal(cs.lapl(f, dx, mode=pbx.MODE_REFERENCE), orc.lapl(f, dx))
    assert np.array_equal(cs.grad(f, dx), orc.grad(f, dx))
    assert np.array_equal(cs.div(v, dx), orc.div(v, dx))
    assert np.array_equal(cs.interp(f), orc.interp(f))
    assert np.array_equal(cs.interp_div(f), orc.interp_div(f))


@pytest.mark.parametrize("shape", [(2048, 8, 8), (8, 2048, 8), (8, 8, 2048), (1000, 12, 10)])
def test_long_lines_reference_bit_exact(shape):
    """BASELINE configs[4]: line lengths up to 2048 in each direction, REFERENCE schedule (these
    bricks have extents that are not multiples of 16, which the FAST schedule does not take)"""
    rng = np.random.default_rng(99)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
    dx = tuple(1.0 / n for n in shape)
    assert np.array_equal(cs.lapl(f, dx, mode=pbx.MODE_REFERENCE), orc.lapl(f, dx))
    assert np.array_equal(cs.grad(f, dx), orc.grad(f, dx))
    assert np.array_equal(cs.div(v, dx), orc.div(v, dx))
    # a FAST request on an unsupported brick is served by the REFERENCE schedule, not refused
    assert np.array_equal(cs.lapl(f, dx, mode=pbx.MODE_FAST), orc.lapl(f, dx))


def test_fast_long_x_lines():
    """x lines of 2048 and 4096 points on the FAST schedule (several warps per line: chunk states
    through shared memory in both the TMA and the generic x kernel)"""
    for shape in ((2048, 16, 16), (4096, 16, 16)):
        rng = np.random.default_rng(5)
        f = np.asfortranarray(rng.uniform(-1, 1, shape))
        dx = tuple(1.0 / n for n in shape)
        assert_fast_close(cs.lapl(f, dx, mode=pbx.MODE_FAST), orc.lapl(f, dx))


@pytest.mark.parametrize("shape", [(32, 16, 48), (64, 64, 64), (128, 32, 64), (16, 512, 16), (2048, 16, 32),
                                   (16, 1024, 16), (32, 16, 2048), (16, 656, 528)])
def test_grad_div_interp_fast_vs_oracle(shape):
    """FAST line operators (chunked first-order recursion) in the reference's stage order"""
    rng = np.random.default_rng(4321)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    v = np.asfortranarray(rng.uniform(-1, 1, shape + (3,)))
    dx = tuple(1.0 / n for n in shape)
    orc.set_threads(8)
    try:
        want = (orc.grad(f, dx), orc.div(v, dx), orc.interp(f), orc.interp_div(f))
    finally:
        orc.set_threads(1)
    cs.set_host_mode(pbx.MODE_FAST)
    try:
        got = (cs.grad(f, dx), cs.div(v, dx), cs.interp(f), cs.interp_div(f))
    finally:
        cs.set_host_mode(pbx.MODE_REFERENCE)
    for g, w in zip(got, want):
        assert np.max(np.abs(g - w)) <= 1e-13 * np.max(np.abs(w))
    # constants are in the null space of every derivative, exactly
    cs.set_host_mode(pbx.MODE_FAST)
    try:
        gc = cs.grad(np.full(shape, 2.8170923), dx)
    finally:
        cs.set_host_mode(pbx.MODE_REFERENCE)
    assert np.max(np.abs(gc)) <= 100 * EPS * max(shape)


def test_fields_golden():
    for tag in "ab":
        f, v, dx = G[f"f3{tag}_f"], G[f"f3{tag}_v"], G[f"f3{tag}_dx"]
        assert np.array_equal(cs.lapl(f, dx, mode=pbx.MODE_REFERENCE), G[f"f3{tag}_lapl"])
        assert np.array_equal(cs.grad(f, dx), G[f"f3{tag}_grad"])
        assert np.array_equal(cs.div(v, dx), G[f"f3{tag}_div"])
        assert np.array_equal(cs.interp(f), G[f"f3{tag}_interp"])
        assert np.array_equal(cs.interp_div(f), G[f"f3{tag}_interpdiv"])
        assert np.array_equal(pbx.compute_lapl.compute_lapl_pointwise(f, dx), G[f"f3{tag}_star"])
    f, dx = G["f3a_f"], G["f3a_dx"]
    assert_fast_close(cs.lapl(f, dx, mode=pbx.MODE_FAST), G["f3a_lapl"])


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 16, 48), (64, 64, 64), (128, 32, 64),
                                   (48, 80, 112), (1024, 16, 16), (16, 512, 16), (16, 16, 512),
                                   (16, 1024, 16), (16, 16, 1024), (32, 2048, 16), (16, 32, 2048),
                                   (16, 528, 656), (64, 1024, 32), (1024, 1024, 16)])
def test_lapl_fast_vs_oracle(shape):
    """S2 inputs: U[-1,1], default_rng(1234), dx = 1/n.  y and z lines of more than
    512 points (BASELINE configs[4]: up to 2048) run as overlapping segments."""
    rng = np.random.default_rng(1234)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    dx = tuple(1.0 / n for n in shape)
    orc.set_threads(8)
    try:
        ref = orc.lapl(f, dx)
    finally:
        orc.set_threads(1)
    assert_fast_close(cs.lapl(f, dx, mode=pbx.MODE_FAST), ref)


def test_lapl_kat_both_modes():
    """tests/lapl/test_lapl.f90:57-132 through the mirror, both schedules"""
    n = 64
    dx = 2 * np.pi / n
    c = (np.arange(n) + 0.5) * dx
    f = np.sin(c)[:, None, None] + np.sin(c)[None, :, None] + np.sin(c)[None, None, :]
    const = np.full((n, n, n), 2.8170923)
    for mode in (pbx.MODE_REFERENCE, pbx.MODE_FAST):
        out = cs.lapl(const, [dx] * 3, mode=mode)
        assert np.sqrt(np.sum(out**2) / n**3) <= 100 * EPS
        out = cs.lapl(f, [dx] * 3, mode=mode)
        rms = np.sqrt(np.mean((out + f) ** 2))
        assert rms == rms and rms <= 1e-9
    # smooth field: the two schedules differ by rounding only
    a = cs.lapl(f, [dx] * 3, mode=pbx.MODE_REFERENCE)
    b = cs.lapl(f, [dx] * 3, mode=pbx.MODE_FAST)
    assert np.max(np.abs(a - b)) <= FAST_TOL * np.max(np.abs(a))


def test_grad_div_3d_kat():
    """tests/grad/test_grad_3d.f90:60-145 and tests/div/test_div_3d.f90:57-144 through the mirror"""
    n = 64
    dx = 2 * np.pi / n
    c, v = (np.arange(n) + 0.5) * dx, np.arange(n) * dx

    def bc(a, ax):
        sh = [1, 1, 1]
        sh[ax] = -1
        return a.reshape(sh)

    const = np.full((n, n, n), 2.8170923)
    assert np.sqrt(np.sum(cs.grad(const, [dx] * 3) ** 2) / n**3 / 3) <= 100 * EPS
    assert np.sqrt(np.sum((cs.interp(const) - const) ** 2) / n**3) <= 100 * EPS
    f = bc(np.sin(c), 0) + bc(np.sin(c), 1) + bc(np.sin(c), 2)
    df = cs.grad(f, [dx] * 3)
    tot = sum(np.sqrt(np.sum((df[..., ax] - bc(np.cos(v), ax)) ** 2) / n) / n / n for ax in range(3))
    assert tot / 3 <= 1e-11
    F = np.empty((n, n, n, 3), order="F")
    for ax in range(3):
        F[..., ax] = np.broadcast_to(bc(np.sin(v), ax), (n, n, n))
    expect = bc(np.cos(c), 0) + bc(np.cos(c), 1) + bc(np.cos(c), 2)
    assert np.sqrt(np.mean((cs.div(F, [dx] * 3) - expect) ** 2)) <= 1e-9


# ------------------------------------------------------------------------------------ 2nd-order star
@pytest.mark.parametrize("shape", [(64, 64, 64), (12, 9, 7), (3, 3, 3), (100, 37, 65), (256, 16, 48)])
def test_star_bit_exact(shape):
    """compute_lapl_pointwise (src/poissbox.f90:84-148), the operator mfmult applies today:
    bit-identical to the oracle through the compute_lapl mirror"""
    from poissbox_b200 import compute_lapl

    rng = np.random.default_rng(77)
    x = np.asfortranarray(rng.uniform(-1, 1, shape))
    dx = (0.13, 0.29 / shape[1], 1.7)
    assert np.array_equal(compute_lapl.compute_lapl_pointwise(x, dx), orc.star(x, dx))


def test_star_kat_and_matmult_operator():
    """test_star.f90's fields on a periodic grid (constant -> 0, eigenfunctions), and the shell
    matrix switching between the two operators"""
    import torch

    n = 64
    h_ = 2 * np.pi / n
    c = (np.arange(n) + 0.5) * h_
    from poissbox_b200 import compute_lapl

    assert np.max(np.abs(compute_lapl.compute_lapl_pointwise(np.full((n, n, n), 1.848), [h_] * 3))) <= 1e-9
    f = np.sin(c)[:, None, None] + np.sin(c)[None, :, None] + np.sin(c)[None, None, :]
    lam = -(4 / h_**2) * np.sin(h_ / 2) ** 2
    assert np.max(np.abs(compute_lapl.compute_lapl_pointwise(f, [h_] * 3) - lam * f)) <= 1e-10
    h = pbx.Handle(n, n, n, (h_,) * 3)
    t = pbx.fortran_to_torch(f)
    assert torch.equal(h.mult(t), h.lapl(t))
    h.operator = 1
    assert torch.equal(h.mult(t), h.star(t))
    assert np.array_equal(pbx.torch_to_fortran(h.star(t)), orc.star(np.asfortranarray(f), [h_] * 3))
    h.close()


# ------------------------------------------------------------------------------------ full-size properties
def test_lapl_full_size_properties():
    """256^3 (BASELINE configs[2]) device-resident: agreement of the two schedules, constants in
    the null space, linearity, symmetry <u, L v> = <L u, v>, and the fused p.Ap."""
    import torch

    n = 256
    h = pbx.Handle(n, n, n, (1.0 / n,) * 3)
    g = torch.Generator(device="cuda").manual_seed(1234)
    u = torch.rand((n, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v = torch.rand((n, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    h.mode = pbx.MODE_FAST
    Lu, Lv = h.lapl(u), h.lapl(v)
    h.mode = pbx.MODE_REFERENCE
    Lu_ref = h.lapl(u)
    scale = Lu_ref.abs().max().item()
    assert (Lu - Lu_ref).abs().max().item() <= FAST_TOL * scale
    h.mode = pbx.MODE_FAST
    assert h.lapl(torch.full_like(u, 2.8170923)).abs().max().item() <= 1e-9 * scale
    Lw = h.lapl(u + 0.5 * v)
    assert (Lw - (Lu + 0.5 * Lv)).abs().max().item() <= 1e-13 * scale
    a, b = torch.dot(u.flatten(), Lv.flatten()).item(), torch.dot(Lu.flatten(), v.flatten()).item()
    assert abs(a - b) <= 1e-12 * (u.norm() * Lv.norm()).item()
    w, dot = h.lapl_dot(u)
    assert torch.equal(w, Lu)
    ref = torch.dot(u.flatten(), Lu.flatten()).item()
    assert abs(dot.item() - ref) <= 1e-13 * abs(ref)
    assert ref < 0   # negative semi-definite
    h.close()


@pytest.mark.parametrize("shape", [(512, 512, 32), (64, 64, 64), (256, 128, 32), (32, 512, 512), (48, 80, 112),
                                   (64, 1024, 32), (32, 16, 2048), (16, 640, 1088), (1024, 64, 32),
                                   (4096, 16, 16), (2048, 1024, 16), (48, 640, 1088)])
def test_tma_and_generic_kernels_bit_identical(shape):
    """the TMA-pipelined persistent kernels and the generic kernels share their arithmetic
    (pbx_fast_common.cuh): same bits, including the fused p.Ap partial sums"""
    import torch

    nx, ny, nz = shape
    dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
    g = torch.Generator(device="cuda").manual_seed(7)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    outs = []
    for no_tma, tma_yz in (("0", "1"), ("1", "0"), ("0", "0")):
        os.environ["PBX_NO_TMA"] = no_tma
        os.environ["PBX_TMA_YZ"] = tma_yz
        try:
            h = pbx.Handle(nx, ny, nz, dx)
        finally:
            os.environ.pop("PBX_NO_TMA", None)
            os.environ.pop("PBX_TMA_YZ", None)
        h.mode = pbx.MODE_FAST
        w, dot = h.lapl_dot(f)
        torch.cuda.synchronize()
        outs.append((w.clone(), dot.clone()))
        h.close()
    # (48, 640, 1088): a brick beyond the L2 with segmented y AND z lines -- the case in which the segmented TMA tiles
    # returned wrong fields in round 2 until their buffers were released behind a proxy fence (test below)
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0])
        # the per-CTA partial sums of p.Ap carry the same bits; the kernel that adds them up (the z pass's own
        # last CTA with 512 threads, or k_reduce with 256) associates them differently
        assert abs(outs[0][1].item() - o[1].item()) <= 1e-14 * abs(o[1].item())


@pytest.mark.parametrize("shape", [(48, 512, 1088), (512, 512, 64)])
def test_tma_tiles_repeatable(shape):
    """Regression test of round 2's defect: tile buffers released to the TMA refill while shared-memory loads of them
    were still queued returned, in 10-30 % of the applies of a large brick with segmented z lines, one 64-row box of the
    NEXT tile (profiles/r2_seg_defect_rootcause.log).  Thirty applies of the z pass with its fused dot into a
    NaN-filled field must all carry the generic kernels' bits."""
    import torch

    nx, ny, nz = shape
    dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
    g = torch.Generator(device="cuda").manual_seed(7)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    os.environ["PBX_NO_TMA"] = "1"
    try:
        hg = pbx.Handle(nx, ny, nz, dx)
    finally:
        os.environ.pop("PBX_NO_TMA", None)
    ref = hg.lapl(f)
    hg.close()
    h = pbx.Handle(nx, ny, nz, dx)
    out = torch.empty_like(ref)
    bad = 0
    for _ in range(30):
        out.fill_(float("nan"))
        h.lapl_dot(f, out)
        torch.cuda.synchronize()
        bad += int(not torch.equal(out, ref))
    h.close()
    assert bad == 0, f"{bad} of 30 applies differ from the generic kernels"
