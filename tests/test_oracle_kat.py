"""Pins the CPU oracle against every in-scope known-answer test of the reference.

Each test restates one reference test program (file:line cited) with its own pass criterion.  The
reference holds no stored vectors; its tests are analytic (constants, sines, polynomials) or
random diagonally-dominant systems from an unseeded `random_number`, for which a seeded generator
with the same construction is substituted (tests/tridiag/test_tdma_utils.f90:12-67).

A second, independent check (`test_*_spectral`) evaluates the same operators through their
Fourier symbols with numpy FFTs, so the oracle is not only checked against smooth inputs.
"""
import numpy as np
import pytest

import oracle_lib as orc

EPS = np.finfo(np.float64).eps


# ---------------------------------------------------------------- tests/tridiag/test_tdma_utils.f90
def tdma_init(n, rng, periodic=False):
    """test_tdma_utils.f90:12-67: random a,b,c,x; diagonal scaled by 10 until dominant; d = A x."""
    a, b, c, x = (rng.random(n) for _ in range(4))
    if not periodic:
        a[0] = 0.0
        c[n - 1] = 0.0
    for i in range(n):
        while b[i] == 0.0:
            b[i] = rng.random()
        while abs(b[i]) < abs(a[i]) + abs(c[i]):
            b[i] = 10 * b[i]
    d = np.empty(n)
    d[0] = b[0] * x[0] + c[0] * x[1]
    if periodic:
        d[0] = a[0] * x[n - 1] + d[0]
    for i in range(1, n - 1):
        d[i] = a[i] * x[i - 1] + b[i] * x[i] + c[i] * x[i + 1]
    d[n - 1] = a[n - 1] * x[n - 2] + b[n - 1] * x[n - 1]
    if periodic:
        d[n - 1] = c[n - 1] * x[0] + d[n - 1]
    return a, b, c, x, d


def utri_residual(b, c, x, d):
    """test_tdma_sweeps.f90:87-119"""
    n = len(d)
    r = d.copy()
    r[: n - 1] -= b[: n - 1] * x[: n - 1] + c[: n - 1] * x[1:]
    r[n - 1] -= b[n - 1] * x[n - 1]
    return np.sqrt(np.sum(r**2) / n), EPS * np.sqrt(np.sum(d**2) / n)


@pytest.mark.parametrize("seed", range(8))
def test_tdma_sweeps(seed):
    """tests/tridiag/test_tdma_sweeps.f90:37-85, n = 128.  The reference tolerance is one epsilon
    of RMS(d) on a single unseeded draw; over several seeded draws we allow 2 epsilon (SURVEY 4
    notes the reference passes with no margin)."""
    n = 128
    a, b, c, x, d = tdma_init(n, np.random.default_rng(100 + seed))
    bp, dp = orc.fwd_sweep(a, b, c, d)
    r, t = utri_residual(bp, c, x, dp)
    assert r <= 2 * t
    dd = np.empty(n)
    dd[: n - 1] = b[: n - 1] * x[: n - 1] + c[: n - 1] * x[1:]
    dd[n - 1] = b[n - 1] * x[n - 1]
    xs = orc.bwd_sweep(b, c, dd)
    r, t = utri_residual(b, c, xs, dd)
    assert r <= 2 * t


@pytest.mark.parametrize("seed", range(8))
def test_tdma(seed):
    """tests/tridiag/test_tdma.f90:40-65: solves the non-periodic system to ~epsilon and, run on
    a periodic system, must NOT reproduce x (:22-24)."""
    n = 128
    a, b, c, x, d = tdma_init(n, np.random.default_rng(200 + seed))
    xs = orc.tdma(a, b, c, d)
    err = np.sqrt(np.sum((x - xs) ** 2) / n)
    assert err <= 2 * EPS * np.sqrt(np.sum(x**2) / n)
    a, b, c, x, d = tdma_init(n, np.random.default_rng(300 + seed), periodic=True)
    xs = orc.tdma(a, b, c, d)
    err = np.sqrt(np.sum((x - xs) ** 2) / n)
    assert err > EPS * np.sqrt(np.sum(x**2) / n)


@pytest.mark.parametrize("seed", range(8))
def test_tdma_periodic(seed):
    """tests/tridiag/test_tdma_periodic.f90:41-63: periodic and non-periodic systems."""
    n = 128
    for periodic in (True, False):
        a, b, c, x, d = tdma_init(n, np.random.default_rng(400 + seed), periodic=periodic)
        b0 = b.copy()
        xs = orc.tdma_periodic(a, b, c, d)
        err = np.sqrt(np.sum((x - xs) ** 2) / n)
        assert err <= 2 * EPS * np.sqrt(np.sum(x**2) / n)
        assert np.array_equal(b, b0)  # tdma_periodic leaves b untouched (tridsol.f90:54-61)


def test_tdma_periodic_dense():
    """independent: against numpy's dense solve of the cyclic matrix"""
    rng = np.random.default_rng(7)
    for n in (5, 16, 128, 257):
        a, b, c, x, d = tdma_init(n, rng, periodic=True)
        A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
        A[0, n - 1] += a[0]
        A[n - 1, 0] += c[n - 1]
        xs = orc.tdma_periodic(a, b, c, d)
        assert np.allclose(xs, np.linalg.solve(A, d), rtol=0, atol=1e-13)


# ---------------------------------------------------------------- tests/coefficients/test_compact.f90
def test_compact_coefficients():
    """tests/coefficients/test_compact.f90:114-163: the scheme constants satisfy the discrete
    relation exactly for polynomials of degree 0..5.  The reference evaluates the relation by hand
    (it never calls the library), in single-precision-literal parameters; here it is evaluated with
    the oracle's own RHS stencil so that the constants the oracle uses are the ones pinned."""
    L, n = np.float64(np.float32(6.28)), 128.0
    dx = L / n
    coef = [np.float64(np.float32(v)) for v in (3.14, 0.817, -7.362, 8.981, -10.22, 0.071)]
    xs = np.arange(4) * dx
    xm = 1.5 * dx
    f = np.zeros(4)
    df = np.zeros(3)
    fi = np.zeros(3)
    tol = 100 * EPS
    for p, m in enumerate(coef):
        f = f + m * xs**p
        df = df + (p * m) * np.array([(xm - dx), xm, (xm + dx)]) ** (p - 1) if p else df
        fi = fi + m * np.array([(xm - dx), xm, (xm + dx)]) ** p
        # derivative: alpha [df] = b (f4 - f1)/(3dx) + a (f3 - f2)/dx
        a_ = 63.0 / 62.0 / dx
        b_ = 17.0 / 62.0 / (3.0 * dx)
        al = 9.0 / 62.0
        rhs = a_ * (f[2] - f[1]) + b_ * (f[3] - f[0])
        assert abs(rhs - (al * df[0] + df[1] + al * df[2])) <= tol * max(1.0, abs(rhs))
        # the oracle's stencil gives the same number at the interior point
        fl = np.concatenate([f, np.zeros(4)])
        assert orc.eval_1d_rhs(a_, b_, -1, -1, fl)[2] == rhs
        # interpolation
        rhs = 0.75 * (f[2] + f[1]) + (1.0 / 20.0) * (f[3] + f[0])
        assert abs(rhs - (0.3 * fi[0] + fi[1] + 0.3 * fi[2])) <= tol * max(1.0, abs(rhs))
        assert orc.eval_1d_rhs(0.75, 1.0 / 20.0, +1, -1, fl)[2] == rhs


# ---------------------------------------------------------------- tests/grad/test_grad_1d.f90, tests/div/test_div_1d.f90
# ---------------------------------------------------------------- tests/coefficients/test_d2dx2.f90, test_star.f90
F32 = lambda v: float(np.float32(v))   # the tests' parameters are default-real literals


def _feq(val, ref, tol):
    d = abs(val - ref)
    return d <= tol * abs(ref) or d <= tol


def test_d2dx2():
    """tests/coefficients/test_d2dx2.f90: lapl_1d_coeffs on constant / linear / quadratic triples,
    scaled, shifted and on 2 dx and dx / 2 grids, tolerance 100 eps (:183-200)"""
    a, b, c, x, dx, shift = (F32(v) for v in (2.718, 1.414, 1.848, 1.618, 0.155, 17.29))
    pts = np.array([x - dx, x, x + dx])
    tol = 100 * EPS

    def ev(f, h):   # :163-175
        co = orc.lapl_1d_coeffs(h)
        return (co[0] * f[0] + co[2] * f[2]) + co[1] * f[1]

    for f, want, spacing in ((np.full(3, c), 0.0, True), (b * pts, 0.0, True), (a * pts**2, 2 * a, False)):
        assert _feq(ev(f, dx) * dx**2, want * dx**2, tol)
        assert _feq(ev(2 * f, dx), 2 * want, tol)
        assert _feq(ev(f / 2, dx) * dx**2, want * dx**2 / 2, tol)
        assert _feq(ev(f + shift, dx) * dx**2, want * dx**2, tol)
        assert _feq(ev(f - shift, dx) * dx**2, want * dx**2, tol)
        if spacing:
            assert _feq(ev(f, 2 * dx) * (2 * dx) ** 2, want * (2 * dx) ** 2, tol)
            assert _feq(ev(f, dx / 2) * (dx / 2) ** 2, want * (dx / 2) ** 2, tol)


def test_star():
    """tests/coefficients/test_star.f90: lapl_star_coeffs + dot_product on 3x3x3 boxes holding a
    constant, a constant-gradient and a quadratic field; tolerance 100 * 1.1 eps (:160-169).  The
    same boxes go through evaluate_laplacian_pointwise (src/poissbox.f90:128-148)."""
    a, b, c, x, dx = (F32(v) for v in (2.718, 1.414, 1.848, 1.618, 0.155))
    pts = np.array([x - dx, x, x + dx])
    tol = 100 * (1.1 * EPS)

    def box(line):   # :45-79
        return line[:, None, None] + line[None, :, None] + line[None, None, :]

    fc = np.full((3, 3, 3), c)
    for f, want in ((fc, 0.0), (box(b * pts), 0.0), (box(a * pts**2), 3 * (2 * a))):
        co = orc.lapl_star_coeffs(dx, dx, dx)
        val = float(np.dot(f.reshape(-1, order="F"), co.reshape(-1, order="F")))
        assert _feq(val * dx**2, want * dx**2, tol)
        val = orc.evaluate_laplacian_pointwise(np.asfortranarray(f), (dx, dx, dx))
        assert _feq(val * dx**2, want * dx**2, tol)
    co = orc.lapl_star_coeffs(0.1, 0.2, 0.4)
    assert np.count_nonzero(co) == 7 and co[1, 1, 1] == ((0 - 2.0 * (1 / 0.1**2)) - 2.0 * (1 / 0.2**2)) - 2.0 * (1 / 0.4**2)


def test_star_periodic_field():
    """compute_lapl_pointwise (src/poissbox.f90:84-126) on a periodic box: constants are in the null
    space, a sine is an eigenfunction with eigenvalue -(4/h^2) sin^2(k h / 2) per direction, and the
    dense 27-term dot product equals the 7-term sum"""
    n = (12, 9, 16)
    h = tuple(2 * np.pi / m for m in n)
    assert np.max(np.abs(orc.star(np.full(n, 2.8170923), h))) <= 1e-10
    xs = [(np.arange(m) + 0.5) * d for m, d in zip(n, h)]
    f = np.sin(xs[0])[:, None, None] + np.sin(2 * xs[1])[None, :, None] + np.cos(3 * xs[2])[None, None, :]
    lam = [-(4 / d**2) * np.sin(k * d / 2) ** 2 for k, d in zip((1, 2, 3), h)]
    want = lam[0] * np.sin(xs[0])[:, None, None] + lam[1] * np.sin(2 * xs[1])[None, :, None] + lam[2] * np.cos(3 * xs[2])[None, None, :]
    got = orc.star(np.asfortranarray(f), h)
    assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))
    rng = np.random.default_rng(2)
    g = np.asfortranarray(rng.uniform(-1, 1, n))
    cx, cy, cz = (1 / d**2 for d in h)
    seven = (cx * (np.roll(g, 1, 0) + np.roll(g, -1, 0)) + cy * (np.roll(g, 1, 1) + np.roll(g, -1, 1))
             + cz * (np.roll(g, 1, 2) + np.roll(g, -1, 2)) - 2 * (cx + cy + cz) * g)
    assert np.max(np.abs(orc.star(g, h) - seven)) <= 1e-12 * np.max(np.abs(seven))


def test_grad_1d_interp_1d():
    """tests/grad/test_grad_1d.f90:53-134 (n = 128, L = 2 pi)"""
    n = 128
    dx = 2 * np.pi / n
    f = np.full(n, 2.8170923)
    assert np.sqrt(np.sum(orc.grad_1d(f, dx) ** 2) / n) <= 100 * EPS
    assert np.sqrt(np.sum((orc.interp_1d(f) - f) ** 2) / n) <= 100 * EPS
    xc = (np.arange(n) + 0.5) * dx
    xv = np.arange(n) * dx
    f = np.sin(xc)
    assert np.sqrt(np.mean((orc.grad_1d(f, dx) - np.cos(xv)) ** 2)) <= 1e-11
    assert np.sqrt(np.mean((orc.interp_1d(f) - np.sin(xv)) ** 2)) <= 1e-11


def test_div_1d_interp_1d_div():
    """tests/div/test_div_1d.f90:53-134"""
    n = 128
    dx = 2 * np.pi / n
    f = np.full(n, 2.8170923)
    assert np.sqrt(np.sum(orc.div_1d(f, dx) ** 2) / n) <= 100 * EPS
    assert np.sqrt(np.sum((orc.interp_1d_div(f) - f) ** 2) / n) <= 100 * EPS
    xc = (np.arange(n) + 0.5) * dx
    xv = np.arange(n) * dx
    f = np.sin(xv)
    assert np.sqrt(np.mean((orc.div_1d(f, dx) - np.cos(xc)) ** 2)) <= 1e-11
    assert np.sqrt(np.mean((orc.interp_1d_div(f) - np.sin(xc)) ** 2)) <= 1e-11


def test_size_mismatch_is_error_7():
    """compact_schemes.f90:177-180, 292-295: `stop 7` on a length mismatch"""
    import ctypes

    f = np.zeros(8)
    g = np.zeros(7)
    dp = ctypes.POINTER(ctypes.c_double)
    assert orc.LIB.orc_grad_1d(8, f.ctypes.data_as(dp), 0.1, 7, g.ctypes.data_as(dp), -1) == 7
    assert orc.LIB.orc_interp_1d(8, f.ctypes.data_as(dp), 7, g.ctypes.data_as(dp), -1) == 7


# ---------------------------------------------------------------- tests/grad/test_grad_3d.f90
def _grid(n):
    dx = 2 * np.pi / n
    c = (np.arange(n) + 0.5) * dx
    v = np.arange(n) * dx
    return dx, c, v


def _bc(a, axis):
    sh = [1, 1, 1]
    sh[axis] = -1
    return a.reshape(sh)


def test_grad_3d():
    """tests/grad/test_grad_3d.f90:60-355 (64^3): constant field, sin x + sin y + sin z, and each
    direction alone, with the test's own normalisation sqrt(sum/nx)/ny/nz (:139-145)."""
    n = 64
    dx, c, v = _grid(n)
    const = np.full((n, n, n), 2.8170923, order="F")
    df = orc.grad(const, [dx] * 3)
    assert np.sqrt(np.sum(df**2) / n**3 / 3) <= 100 * EPS
    assert np.sqrt(np.sum((orc.interp(const) - const) ** 2) / n**3) <= 100 * EPS

    def check(use):
        f = np.zeros((n, n, n), order="F")
        for ax in range(3):
            if use[ax]:
                f = f + _bc(np.sin(c), ax)
        df = orc.grad(f, [dx] * 3)
        tot = 0.0
        for ax in range(3):
            expect = _bc(np.cos(v), ax) if use[ax] else 0.0
            tot += np.sqrt(np.sum((df[..., ax] - expect) ** 2) / n) / n / n
        assert tot / 3 <= 1e-11

    check((1, 1, 1))
    check((1, 0, 0))
    check((0, 1, 0))
    check((0, 0, 1))


# ---------------------------------------------------------------- tests/div/test_div_3d.f90
def test_div_3d():
    """tests/div/test_div_3d.f90:57-144"""
    n = 64
    dx, c, v = _grid(n)
    const = np.full((n, n, n, 3), 2.8170923, order="F")
    assert np.sqrt(np.sum(orc.div(const, [dx] * 3) ** 2) / n**3) <= 100 * EPS
    assert np.sqrt(np.sum((orc.interp_div(const[..., 0]) - const[..., 0]) ** 2) / n**3) <= 100 * EPS
    F = np.empty((n, n, n, 3), order="F")
    for ax in range(3):
        F[..., ax] = np.broadcast_to(_bc(np.sin(v), ax), (n, n, n))
    expect = _bc(np.cos(c), 0) + _bc(np.cos(c), 1) + _bc(np.cos(c), 2)
    assert np.sqrt(np.mean((orc.div(F, [dx] * 3) - expect) ** 2)) <= 1e-9


# ---------------------------------------------------------------- tests/lapl/test_lapl.f90
def test_lapl():
    """tests/lapl/test_lapl.f90:57-132"""
    n = 64
    dx, c, v = _grid(n)
    const = np.full((n, n, n), 2.8170923, order="F")
    assert np.sqrt(np.sum(orc.lapl(const, [dx] * 3) ** 2) / n**3) <= 100 * EPS
    f = _bc(np.sin(c), 0) + _bc(np.sin(c), 1) + _bc(np.sin(c), 2)
    out = orc.lapl(f, [dx] * 3)
    rms = np.sqrt(np.mean((out + f) ** 2))
    assert rms == rms and rms <= 1e-9
    assert abs(rms - 3.73e-10) < 0.1e-10  # SURVEY 4 [derived] magnitude


# ---------------------------------------------------------------- independent spectral evaluation
def _symbols(n, h):
    th = 2 * np.pi * np.fft.fftfreq(n)
    a, b, al = 63.0 / 62.0 / h, 17.0 / 62.0 / (3.0 * h), 9.0 / 62.0
    # cell -> vertex (stagger -1): vertex i sits half a cell left of cell i
    kd = 1j * (2 * a * np.sin(th / 2) + 2 * b * np.sin(3 * th / 2)) / (1 + 2 * al * np.cos(th))
    ti = (1.5 * np.cos(th / 2) + 0.1 * np.cos(3 * th / 2)) / (1 + 0.6 * np.cos(th))
    ph = np.exp(-0.5j * th)  # half-cell shift of the stagger
    return kd, ti, ph


def test_lapl_spectral():
    """lapl symbol = -sum_d k'_d^2 prod_{e != d} T_e^2, random field, non-cubic"""
    rng = np.random.default_rng(1234)
    nx, ny, nz = 32, 48, 20
    dx = [1.0 / nx, 0.7 / ny, 1.3 / nz]
    f = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
    kd = []
    ti = []
    for n, h in zip((nx, ny, nz), dx):
        k, t, _ = _symbols(n, h)
        kd.append((k * np.conj(k)).real)
        ti.append(t * t)
    KX, KY, KZ = np.meshgrid(kd[0], kd[1], kd[2], indexing="ij")
    TX, TY, TZ = np.meshgrid(ti[0], ti[1], ti[2], indexing="ij")
    sym = -(KX * TY * TZ + TX * KY * TZ + TX * TY * KZ)
    ref = np.fft.ifftn(np.fft.fftn(f) * sym).real
    out = orc.lapl(f, dx)
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_1d_spectral():
    rng = np.random.default_rng(5)
    n, h = 96, 0.37
    f = rng.uniform(-1, 1, n)
    kd, ti, ph = _symbols(n, h)
    F = np.fft.fft(f)
    for out, sym in (
        (orc.grad_1d(f, h), kd * ph),
        (orc.div_1d(f, h), kd * np.conj(ph)),
        (orc.interp_1d(f), ti * ph),
        (orc.interp_1d_div(f), ti * np.conj(ph)),
    ):
        ref = np.fft.ifft(F * sym).real
        assert np.max(np.abs(out - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


def test_threads_bit_identical():
    rng = np.random.default_rng(11)
    f = np.asfortranarray(rng.uniform(-1, 1, (24, 16, 20)))
    a = orc.lapl(f, [0.1, 0.2, 0.3])
    orc.set_threads(4)
    try:
        b = orc.lapl(f, [0.1, 0.2, 0.3])
    finally:
        orc.set_threads(1)
    assert np.array_equal(a, b)


# ---------------------------------------------------------------- CG (parity unpinned, see oracle header)
def test_cg_converges_manufactured():
    """demo recipe (src/example.f90:70-72): random x_true, b = A x_true, x0 = 0, constant null
    space.  CG must converge to rtol and recover x_true up to the operator's null space."""
    n = 16
    rng = np.random.default_rng(1234)
    xt = np.asfortranarray(rng.uniform(-1, 1, (n, n, n)))
    dx = [1.0 / n] * 3
    b = orc.lapl(xt, dx)
    x, its, rnorm, reason, hist = orc.cg_solve(b, dx, rtol=1e-8)
    assert reason == 2 and 0 < its < 200
    assert rnorm <= 1e-8 * hist[0]
    r = orc.lapl(x, dx) - b
    assert np.linalg.norm(r) <= 1e-7 * np.linalg.norm(b)
