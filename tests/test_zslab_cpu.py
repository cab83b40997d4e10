"""CPU tests of the N>1 (z-slab) path: the correction tables the C library builds, and the
moments -> exchange -> correction protocol, first in one process and then across two processes with
torch.distributed (gloo), against a dense periodic evaluation of the z pass."""
import os
import subprocess
import sys

import numpy as np
import pytest

import zslab_model as zm

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("P,nzl,dz", [(2, 64, 1 / 128), (8, 64, 1 / 512), (4, 128, 2 * np.pi / 512), (3, 80, 0.01)])
def test_slab_protocol_single_process(P, nzl, dz):
    rng = np.random.default_rng(P * 1000 + nzl)
    nl = 5
    c = rng.uniform(-1, 1, (P * nzl, nl))
    d = rng.uniform(-1, 1, (P * nzl, nl))
    T = zm.tables(nzl, dz)
    assert T["R"][0] <= 8 and T["R"][1] <= 8 and T["R"][0] >= 4
    mom = [zm.moments(T, c[p * nzl:(p + 1) * nzl], d[p * nzl:(p + 1) * nzl]) for p in range(P)]
    out = np.empty_like(c)
    for p in range(P):
        lo, up = (p - 1) % P, (p + 1) % P
        loc = zm.local_open(c[p * nzl:(p + 1) * nzl], d[p * nzl:(p + 1) * nzl], dz)
        m_a = mom[lo][0] + mom[p][2]      # lower's send_up + my self_a
        m_b = mom[up][1] + mom[p][3]      # upper's send_dn + my self_b
        out[p * nzl:(p + 1) * nzl] = zm.correct(T, loc, m_a, m_b)
    truth = zm.periodic_truth(c, d, dz)
    assert np.max(np.abs(out - truth)) <= 5e-15 * np.max(np.abs(truth))


@pytest.mark.parametrize("P,nzl,dz", [(2, 64, 1 / 128), (8, 64, 1 / 512), (4, 128, 2 * np.pi / 512), (3, 80, 0.01)])
def test_slab_protocol_native_inputs(P, nzl, dz):
    """the protocol the CUDA path uses: states and halos rebuilt from 9 numbers per direction"""
    rng = np.random.default_rng(P * 77 + nzl)
    c = rng.uniform(-1, 1, (P * nzl, 4))
    d = rng.uniform(-1, 1, (P * nzl, 4))
    msgs = [zm.boundary_messages(c[p * nzl:(p + 1) * nzl], d[p * nzl:(p + 1) * nzl], dz) for p in range(P)]
    out = np.empty_like(c)
    for p in range(P):
        out[p * nzl:(p + 1) * nzl] = zm.slab_zpass(c[p * nzl:(p + 1) * nzl], d[p * nzl:(p + 1) * nzl], dz,
                                                   msgs[(p - 1) % P][0], msgs[(p + 1) % P][1])
    truth = zm.periodic_truth(c, d, dz)
    assert np.max(np.abs(out - truth)) <= 5e-15 * np.max(np.abs(truth))


def test_too_thin_slab_rejected():
    import poissbox_b200 as pbx

    with pytest.raises(pbx.PbxError) as e:
        zm.tables(48, 0.01)
    assert e.value.code == 4


def test_slab_protocol_gloo_world2():
    """two processes, gloo: each owns one slab and exchanges its moments with send/recv"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "zslab_gloo_worker.py"), str(r), "2"],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ZSLAB_OK" in o, o


def test_slab_kernels_gloo_world2():
    """two processes, gloo, and the library's own slab kernels (on the CPU kernel-logic harness):
    Laplacian, grad, div, interp and star with the exchange owned by the host
    (pbx_slab_get_messages / send-recv / pbx_slab_put_messages), plus the CG's all-reduced dot"""
    import emu_lib

    emu_lib.load()   # build once before the ranks start
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "emu_gloo_worker.py"), str(r), "2"],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "EMU_GLOO_OK" in o, o


@pytest.mark.parametrize("world,port,fuse", [(2, 29621, "0"), (3, 29623, "0"), (2, 29625, "1")])
def test_peer_boards(world, port, fuse):
    """`world` processes (gloo rendezvous) sharing their receive buffers: the library's own multi-rank
    paths -- peer stores of the boundary messages, the neighbour flag barrier, the all-reduce fused
    into the CG's reduction kernel (pbx_slab_link_peers) -- on the CPU kernel-logic harness, with no
    host-side exchange at all: Laplacian, grad, div, interp, star, dot and the distributed CG against
    one handle on the whole brick (same iteration counts, same bits of the sums on every rank).
    fuse = "1": PBX_FUSE_TAIL -- the z pass and the residual update reduce their partial sums, all-reduce
    them and run the CG's scalar step in their own last CTA (five launches per iteration)"""
    import emu_lib

    emu_lib.load()   # build once before the ranks start
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PBX_FUSE_TAIL=fuse)
    tag = f"{os.getpid()}w{world}f{fuse}"
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "emu_peer_worker.py"), str(r), str(world), tag],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "EMU_PEER_OK" in o, o
