"""CPU: the oracle reproduces the committed golden vectors bit for bit (tests/golden/make_golden.py)."""
import os

import numpy as np

import oracle_lib as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))


def test_tridsol_golden():
    for n in (33, 128):
        for per in (0, 1):
            k = f"tri_n{n}_p{per}"
            a, b, c, d = G[k + "_abcd"]
            bp, dp = orc.fwd_sweep(a, b, c, d)
            assert np.array_equal(np.stack([bp, dp]), G[k + "_fwd"])
            assert np.array_equal(orc.bwd_sweep(b, c, d), G[k + "_bwd"])
            assert np.array_equal(orc.tdma(a, b, c, d), G[k + "_tdma"])
            assert np.array_equal(orc.tdma_periodic(a, b, c, d), G[k + "_tdmap"])


def test_lines_golden():
    for n in (37, 128):
        f = G[f"l1_n{n}_f"]
        dx = 1.0 / n
        assert np.array_equal(orc.grad_1d(f, dx), G[f"l1_n{n}_grad"])
        assert np.array_equal(orc.div_1d(f, dx), G[f"l1_n{n}_div"])
        assert np.array_equal(orc.interp_1d(f), G[f"l1_n{n}_interp"])
        assert np.array_equal(orc.interp_1d_div(f), G[f"l1_n{n}_interpdiv"])


def test_fields_golden():
    for tag in "ab":
        f, v, dx = G[f"f3{tag}_f"], G[f"f3{tag}_v"], G[f"f3{tag}_dx"]
        assert np.array_equal(orc.lapl(f, dx), G[f"f3{tag}_lapl"])
        assert np.array_equal(orc.grad(f, dx), G[f"f3{tag}_grad"])
        assert np.array_equal(orc.div(v, dx), G[f"f3{tag}_div"])
        assert np.array_equal(orc.interp(f), G[f"f3{tag}_interp"])
        assert np.array_equal(orc.interp_div(f), G[f"f3{tag}_interpdiv"])


def test_cg_golden():
    b = G["cg16_b"]
    x, its, rnorm, reason, hist = orc.cg_solve(b, (1 / 16,) * 3, rtol=1e-8)
    assert (its, reason) == (int(G["cg16_meta"][0]), int(G["cg16_meta"][1]))
    assert np.array_equal(hist, G["cg16_hist"])
    assert np.array_equal(x, G["cg16_x"])


def test_star_golden():
    for tag in "ab":
        assert np.array_equal(orc.star(G[f"f3{tag}_f"], G[f"f3{tag}_dx"]), G[f"f3{tag}_star"])
    x, its, rnorm, reason, hist = orc.cg_solve(G["cgstar16_b"], (1 / 16,) * 3, rtol=1e-8, op=1)
    assert (its, reason) == (int(G["cgstar16_meta"][0]), int(G["cgstar16_meta"][1]))
    assert np.array_equal(hist, G["cgstar16_hist"]) and np.array_equal(x, G["cgstar16_x"])
