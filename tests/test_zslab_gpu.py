"""GPU: the z-slab decomposition emulated in one process on one GPU (P phase-driven slab handles,
exchange by device copies) must reproduce the single-handle periodic Laplacian."""
import os

import numpy as np
import pytest

import poissbox_b200 as pbx
from poissbox_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,P", [((64, 32, 128), 2), ((32, 32, 512), 8), ((48, 16, 256), 2),
                                     ((512, 16, 192), 3), ((16, 512, 128), 2), ((32, 1024, 128), 2)])
@pytest.mark.parametrize("no_tma", ["0", "1"])
def test_slabs_match_single_brick(shape, P, no_tma):
    import torch

    nx, ny, nz = shape
    nzl = nz // P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    g = torch.Generator(device="cuda").manual_seed(5)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    os.environ["PBX_NO_TMA"] = no_tma
    try:
        whole = pbx.Handle(nx, ny, nz, dx)
        slabs = [pbx.Handle(nx, ny, nzl, dx, slab=(r, P)) for r in range(P)]
    finally:
        os.environ.pop("PBX_NO_TMA", None)
    ref = whole.lapl(f)
    parts = [f[r * nzl:(r + 1) * nzl].contiguous() for r in range(P)]
    for h, part in zip(slabs, parts):
        h.slab_phase1(part)
    pbx.Handle.slab_exchange_local(slabs)
    out = torch.cat([h.slab_phase2() for h in slabs], dim=0)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item()
    assert err <= 1e-13 * scale, err / scale
    for h in slabs + [whole]:
        h.close()


@pytest.mark.parametrize("shape,P", [((64, 32, 128), 2), ((32, 48, 256), 4), ((16, 1024, 128), 2)])
def test_slab_grad_div_interp_match_single_brick(shape, P):
    """grad / div / interp on P slabs (phase 1, exchange of three numbers per z line, operator and
    neighbour, phase 2) against the whole periodic brick"""
    import torch

    from poissbox_b200 import _lib

    nx, ny, nz = shape
    nzl = nz // P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    g = torch.Generator(device="cuda").manual_seed(9)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    v = torch.rand((3, nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    whole = pbx.Handle(nx, ny, nz, dx)
    slabs = [pbx.Handle(nx, ny, nzl, dx, slab=(r, P)) for r in range(P)]
    want = {_lib.OP_GRAD: whole.grad(f), _lib.OP_DIV: whole.div(v), _lib.OP_INTERP: whole.interp(f),
            _lib.OP_INTERP_DIV: whole.interp(f, +1)}
    for op, w in want.items():
        zdim = 1 if op in (_lib.OP_GRAD,) else 0
        parts = [(v[:, r * nzl:(r + 1) * nzl] if op == _lib.OP_DIV else f[r * nzl:(r + 1) * nzl]).contiguous()
                 for r in range(P)]
        for h, part in zip(slabs, parts):
            h.slab_op_phase1(op, part)
        pbx.Handle.slab_exchange_local(slabs)
        out = torch.cat([h.slab_op_phase2(op, part) for h, part in zip(slabs, parts)], dim=zdim)
        torch.cuda.synchronize()
        err = (out - w).abs().max().item() / w.abs().max().item()
        assert err <= 1e-13, (op, err)
    for h in slabs + [whole]:
        h.close()


def test_slab_host_owned_exchange():
    """the exchange owned by the host: pbx_slab_get_messages / pbx_slab_put_messages instead of
    pbx_slab_exchange_local (what an MPI host does between phase 1 and phase 2)"""
    import torch

    nx, ny, nz, P = 64, 32, 128, 2
    nzl = nz // P
    dx = (1.0 / nx, 0.7 / ny, 1.3 / nz)
    g = torch.Generator(device="cuda").manual_seed(6)
    f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    whole = pbx.Handle(nx, ny, nz, dx)
    ref = whole.lapl(f)
    slabs = [pbx.Handle(nx, ny, nzl, dx, slab=(r, P)) for r in range(P)]
    for r, h in enumerate(slabs):
        h.slab_phase1(f[r * nzl:(r + 1) * nzl].contiguous())
    msgs = [h.slab_get_messages() for h in slabs]          # (to rank+1, to rank-1)
    for r, h in enumerate(slabs):
        h.slab_put_messages(msgs[(r - 1) % P][0], msgs[(r + 1) % P][1])
    out = torch.cat([h.slab_phase2() for h in slabs], dim=0)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 1e-13 * ref.abs().max().item()
    for h in slabs + [whole]:
        h.close()


def test_slab_needs_64_to_512_planes():
    for nzl in (48, 528):
        with pytest.raises(pbx.PbxError) as e:
            pbx.Handle(32, 32, nzl, (1, 1, 1), slab=(0, 2))
        assert e.value.code == 4


@pytest.mark.parametrize("P", [2, 4])
def test_peer_boards_one_gpu(P):
    """P slab handles of one process on one GPU, each on its own stream and host thread, linked by
    pbx_slab_link_peers: boundary messages by direct stores, flag barrier, all-reduce inside the CG's
    reduction kernel -- against one handle on the whole brick.  (A rank's barrier kernel spins on the
    device until its neighbours' boundary sweeps have run, so the ranks' streams must not share a
    hardware queue: P + 2 streams stay below the default CUDA_DEVICE_MAX_CONNECTIONS of 8.)"""
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

    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
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
