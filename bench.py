#!/usr/bin/env python
"""bench.py -- headline benchmark of poissbox-b200.

Metric (BASELINE.json): GDoF/s of one compact-Laplacian apply (the MATSHELL MatMult of the CG
Poisson solve) on a 512^3 fp64 periodic box, plus the CG time-to-1e-8 on the same grid.

  step      one application of the compact Laplacian to one 512^3 field (N>1: each rank owns a
            z-slab of the 512^2 x 512 box ... see --scaling)
  value     whole-job GDoF/s with the field resident in HBM (CUDA events, max over ranks)
  e2e       the same step through the host-pointer C-ABI call a Fortran caller makes
            (pbx_lapl_host): pinned host buffers, H2D + D2H inside the timed region
  roofline  the dominant kernel (y pass, 32 B/DoF algorithmic) against the measured HBM peak
  cpu_baseline  the CPU oracle (op-for-op port of the reference Fortran; the reference itself
            cannot be built here: no Fortran, MPI or PETSc) on a bounded sample, 1 core

`--impl reference` times that CPU port with all host threads instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the slab exchange is two 16 MiB neighbour messages per rank: give NCCL's send/recv path more
# channels than its default of a few per peer (must be set before NCCL initialises)
os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
os.environ.setdefault("NCCL_NCHANNELS_PER_NET_PEER", "16")

METRIC = "GDoF/s of compact Laplacian apply; CG time-to-1e-8 at 512^3 fp64, 1-8 B200"
ALG_BYTES_MATMULT = 80.0     # B/DoF, one sweep per axis
ALG_BYTES_PASS = {"x": 24.0, "y": 32.0, "z": 24.0}
FALLBACK_HBM_GBS = 6650.0    # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # `--grid`: the spelling to use under torchrun, whose own parser rejects `--n` as an ambiguous abbreviation
    ap.add_argument("--n", "--grid", dest="n", type=int, default=512, help="global grid is n^3")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="matmult", choices=["matmult", "config5"],
                    help="config5 = BASELINE configs[4]: standalone batched tridsol + compact grad / div / Laplacian sweeps, "
                         "line lengths 64-2048 along x, y, z (one GPU; one JSON line)")
    ap.add_argument("--quick", action="store_true", help="headline legs only (no 256^3 / S3 solves, no CG through the host call)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check that precedes the timed region")
    ap.add_argument("--cg-rtol", type=float, default=1e-8)
    ap.add_argument("--cg-maxit", type=int, default=20000)
    return ap.parse_args()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs (one streaming
    nvidia-smi process, a line every 50 ms)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self, t_lo=None, t_hi=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.th:
            self.th.join(timeout=5)
        rows = [r for t, r in self.rows if (t_lo is None or t >= t_lo) and (t_hi is None or t <= t_hi)]
        if len(rows) < 3:          # a very short timed region: keep every sample taken under load
            rows = [r for _, r in self.rows]

        def num(v):
            try:
                return float(v)
            except Exception:
                return None

        sm = sorted(v for v in (num(r[0]) for r in rows if r) if v is not None)
        mx = [v for v in (num(r[1]) for r in rows if len(r) > 1) if v is not None]
        pw = [v for v in (num(r[2]) for r in rows if len(r) > 2) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# CPU side (oracle = op-for-op port of the reference; test infrastructure, used here only as the
# reported CPU baseline / the --impl reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_lapl_rate(n, threads, reps):
    """GDoF/s of the oracle Laplacian on an n^3 S2 field with `threads` host threads"""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    rng = np.random.default_rng(1234)
    f = np.asfortranarray(rng.uniform(-1, 1, (n, n, n)))
    dx = (1.0 / n,) * 3
    orc.set_threads(threads)
    try:
        t0 = time.perf_counter()
        for _ in range(reps):
            orc.lapl(f, dx)
        dt = (time.perf_counter() - t0) / reps
    finally:
        orc.set_threads(1)
    return n**3 / dt / 1e9, dt


def cpu_baseline_block():
    # bounded sample: one 256^3 brick of the 512^3 workload (same operator, same line layout;
    # about 16 s on one core), after a 64^3 warm-up
    cpu_lapl_rate(64, 1, 1)
    n = 256
    rate, dt = cpu_lapl_rate(n, 1, 1)
    return {"value": rate, "unit": "GDoF/s", "cores": 1, "kind": "port",
            "sample": f"one compact-Laplacian apply on a {n}^3 S2 brick ({dt:.1f} s), CPU oracle "
                      "(C restatement of the reference Fortran, gcc -O2 no-FMA), 1 thread = the "
                      "reference's serial behaviour"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # calibrate so that (steps + warmup) samples finish in about two minutes
    rate64, _ = cpu_lapl_rate(64, cores, 1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = 64
    cap = int(os.environ.get("PBX_BENCH_REF_N", "0"))   # tests: cap the sample brick (tests/test_bench_contract.py)
    for cand in (96, 128, 192, 256, 384, 512):
        if cand**3 / (rate64 * 1e9) <= budget and cand <= args.n and (cap <= 0 or cand <= cap):
            n = cand
    for _ in range(args.warmup):
        cpu_lapl_rate(n, cores, 1)
    rate, dt = cpu_lapl_rate(n, cores, max(1, args.steps))
    sample = (f"each step = one compact-Laplacian apply on a {n}^3 S2 brick of the {args.n}^3 workload, "
              f"CPU oracle (C port of the reference Fortran; the Fortran+PETSc reference cannot be built "
              f"in this image), {cores} host threads over lines")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"compact Laplacian apply, {args.n}^3 fp64 periodic box, S2 random field",
                   "grid": [args.n] * 3, "sample_grid": [n] * 3},
        "cpu_baseline": {"value": rate, "unit": "GDoF/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def oracle_parity(pbx, torch, dist, world, rank, local, comm):
    """Before any timing: the (distributed) operators of THIS run against the CPU oracle on a small brick,
    64 x 64 x (64 * N) S2 field (every rank owns a 64-plane slab: the thin-slab wrap cases of
    src/compact_schemes.f90:356-370 across every rank boundary), plus the CG iteration count on b = A x_true
    (S3, rtol 1e-5, the north star's +-1).  The oracle is the checker here, never the thing measured."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    nx = ny = 64
    nzl = 64
    nz = nzl * world
    dx = (1.0 / nx, 1.0 / ny, 1.0 / nz)
    rng = np.random.default_rng(1234)
    f = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz)))
    v = np.asfortranarray(rng.uniform(-1, 1, (nx, ny, nz, 3)))
    sl = slice(rank * nzl, (rank + 1) * nzl)
    orc.set_threads(max(1, (os.cpu_count() or 1) // world))
    try:
        want = {"lapl": orc.lapl(f, dx), "grad": orc.grad(f, dx), "div": orc.div(v, dx), "interp": orc.interp(f, -1)}
        its_o = None
        if rank == 0:
            _, its_o, _, why_o, _ = orc.cg_solve(want["lapl"], dx, rtol=1e-5)
    finally:
        orc.set_threads(1)
    h = pbx.Handle(nx, ny, nzl, dx, device=local, comm=comm)
    h.use_current_stream()
    dev = torch.device("cuda", local)
    t = lambda a: pbx.fortran_to_torch(a, device=dev)   # f(i,j,k[,c]) -> (c,) k, j, i: the Fortran memory layout
    fl = t(f[:, :, sl])
    got = {"lapl": pbx.torch_to_fortran(h.lapl(fl)), "grad": pbx.torch_to_fortran(h.grad(fl)),
           "div": pbx.torch_to_fortran(h.div(t(v[:, :, sl, :]))), "interp": pbx.torch_to_fortran(h.interp(fl, -1))}
    errs = {}
    for k in want:
        w = want[k][:, :, sl]
        errs[k] = float(np.max(np.abs(got[k] - w)) / np.max(np.abs(want[k])))
    x, its, _, why, _ = h.cg_solve(t(want["lapl"][:, :, sl]), rtol=1e-5)
    torch.cuda.synchronize()
    h.close()
    e = torch.tensor([errs[k] for k in ("lapl", "grad", "div", "interp")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    e = e.tolist()
    out = {"grid": [nx, ny, nz], "slab_planes": nzl, "field": "S2 U[-1,1], numpy default_rng(1234)",
           "against": "CPU oracle (C restatement of src/compact_schemes.f90), same inputs, outside the timed region",
           "max_abs_err_over_max_ref": {"lapl": e[0], "grad": e[1], "div": e[2], "interp": e[3]},
           "n": nx * ny * nz, "tolerance": 1e-12,
           "cg_its": {"ours": int(its), "oracle": its_o, "rtol": 1e-5, "rhs": "S3: b = A x_true"}, "ok": None}
    if rank == 0:
        out["ok"] = bool(max(e) <= 1e-12 and why == 2 and why_o == 2 and abs(its - its_o) <= 1)
    return out


def cg_case(pbx, torch, n, kind, rtol, local, maxit=20000):
    """one device-resident CG solve on an n^3 box, one GPU: S3 (x_true ~ U[-1,1], L = 1, b = A x_true: the demo's
    recipe, src/example.f90:70-72,180-183) or S4 (manufactured smooth u = exp(sin x + sin y + sin z), L = 2 pi)"""
    import math

    dev = torch.device("cuda", local)
    if kind == "S3":
        hh = 1.0 / n
        g = torch.Generator(device=dev).manual_seed(1234)
        u = torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    else:
        hh = 2 * math.pi / n
        c = (torch.arange(n, dtype=torch.float64, device=dev) + 0.5) * hh
        u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
    h = pbx.Handle(n, n, n, (hh,) * 3, device=local)
    h.use_current_stream()
    b = h.lapl(u)
    del u
    x = h.empty()
    h.cg_solve(b, x, rtol=rtol, maxit=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, its, rnorm, reason, hist = h.cg_solve(b, x, rtol=rtol, maxit=maxit)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = h.lapl(x) - b
    true_rel = float((r.norm() / b.norm()).item())
    h.close()
    return {"grid": [n] * 3, "rhs": kind, "rtol": rtol, "its": int(its), "reason": int(reason), "time_s": dt,
            "ms_per_it": dt / max(1, its) * 1e3, "true_residual_rel": true_rel,
            "frac_of_hbm_peak": 152.0 * float(n) ** 3 * its / dt / 1e9 / hbm_peak()[0]}


def cg_vs_oracle(pbx, n, rtol):
    """the north star's iteration criterion on the full-spectrum problem: the library's CG (through the host-pointer
    call) and the oracle's CG (all host threads) on the SAME b = A x_true, S3, n^3"""
    import ctypes

    import numpy as np

    from poissbox_b200 import _lib

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc

    rng = np.random.default_rng(1234)
    xt = np.asfortranarray(rng.uniform(-1, 1, (n, n, n)))
    dx = (1.0 / n,) * 3
    orc.set_threads(os.cpu_count() or 1)
    try:
        b = orc.lapl(xt, dx)
        t0 = time.perf_counter()
        xo, ito, _, whyo, _ = orc.cg_solve(b, dx, rtol=rtol)
        t_orc = time.perf_counter() - t0
    finally:
        orc.set_threads(1)
    x = np.zeros_like(b, order="F")
    its, why, rn = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    t0 = time.perf_counter()
    pbx.check(pbx.LIB.pbx_cg_solve_host(n, n, n, _lib._d3(*dx), b.ctypes.data_as(_lib._dp), x.ctypes.data_as(_lib._dp),
                                        rtol, 1e-50, 10000, pbx.MODE_FAST, ctypes.byref(its), ctypes.byref(rn),
                                        ctypes.byref(why), None, 0))
    t_gpu = time.perf_counter() - t0
    return {"grid": [n] * 3, "rhs": "S3: b = A x_true from the oracle", "rtol": rtol, "its_ours": its.value,
            "its_oracle": int(ito), "reasons": [why.value, int(whyo)],
            "x_rel_diff": float(np.linalg.norm(x - xo) / np.linalg.norm(xo)),
            "oracle_cg_s": t_orc, "ours_host_call_s": t_gpu, "oracle_threads": os.cpu_count() or 1,
            "within_one": bool(abs(its.value - int(ito)) <= 1)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import poissbox_b200 as pbx
    from poissbox_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        import ctypes

        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            pbx.check(pbx.LIB.pbx_comm_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.to(dev)
        dist.broadcast(idbuf, 0)
        raw = (ctypes.c_ubyte * 128)(*idbuf.cpu().tolist())
        c = ctypes.c_void_p()
        pbx.check(pbx.LIB.pbx_comm_init_rank(raw, world, rank, local, ctypes.byref(c)))
        comm = c.value

    parity = None
    if not args.no_parity:
        parity = oracle_parity(pbx, torch, dist, world, rank, local, comm)
        if rank == 0 and not parity["ok"]:
            print(json.dumps({"error": "oracle parity failed before the timed region", "parity": parity}), flush=True)
            raise SystemExit(1)

    n = args.n
    if n % world or (n // world) % 16:
        raise SystemExit("n / gpus must be a multiple of 16")
    nzl = n // world                       # strong scaling: the 512^3 box is z-slab partitioned
    dx = (1.0 / n,) * 3
    h = pbx.Handle(n, n, nzl, dx, device=local, comm=comm)
    h.use_current_stream()
    h.mode = pbx.MODE_FAST

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # S2 input: U[-1,1] (generated on the device: at 512^3 the field is 1 GiB)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    f = torch.rand((nzl, n, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    out = h.empty()
    ndof_total = float(n) ** 3

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up; long enough (~0.3 s) for the clock sampler to see the load the timed region runs
    # under.  The count is the same on every rank: with N>1 each apply contains collectives.
    nw = max(3, args.warmup, int(0.3 / (2.2e-3 / world)))
    for i in range(nw):
        h.lapl(f, out)
        if i % 8 == 7:
            torch.cuda.synchronize()
    barrier()
    t_lo = time.perf_counter() - 0.25
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        h.lapl(f, out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launches - l0
    clocks = sampler.stop(t_lo, time.perf_counter() + 0.05) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = ndof_total / (ms_per_step * 1e-3) / 1e9

    # per-kernel durations (CUDA events between the launches) for the roofline of the dominant kernel
    peak, peak_src = hbm_peak()
    roof = None
    if world == 1:
        pm = h.lapl_profile(f, out, reps=max(3, min(10, args.steps)))
        names = ("x", "y", "z")
        per = {k: {"ms": pm[i], "alg_bytes_per_dof": ALG_BYTES_PASS[k],
                   "achieved_GBs": ALG_BYTES_PASS[k] * ndof_total / (pm[i] * 1e-3) / 1e9}
               for i, k in enumerate(names)}
        dom = max(names, key=lambda k: per[k]["ms"])
        # DRAM bytes per launch and the kernel's name from the newest dated `ncu --set full` summary under
        # profiles/ (r<round>_ncu_full_<n>.json, written from the .ncu-rep of that round's capture)
        traffic, kname, tsrc = None, {"x": "x_tma_kernel", "y": "yz_tma_kernel (y pass)", "z": "yz_tma_kernel (z pass)"}[dom], None
        import glob
        import re

        caps = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_full_{n}.json")),
                      key=lambda q: int(re.search(r"r(\d+)_ncu", os.path.basename(q)).group(1)))
        if caps:
            try:
                cap = json.load(open(caps[-1]))
                for kk in cap["kernels"]:
                    if kk["pass"] == dom:
                        traffic, kname, tsrc = kk["dram_bytes_per_launch"], kk["kernel"], os.path.basename(caps[-1])
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": kname, "achieved": per[dom]["achieved_GBs"],
                "peak": peak, "unit": "GB/s", "frac": per[dom]["achieved_GBs"] / peak, "traffic": traffic, "traffic_source": tsrc,
                "peak_source": peak_src, "passes": per,
                "matmult": {"alg_bytes_per_dof": ALG_BYTES_MATMULT,
                            "achieved_GBs": ALG_BYTES_MATMULT * value,
                            "frac": ALG_BYTES_MATMULT * value / peak}}
    else:
        roof = {"bound": "hbm", "kernel": "matmult (x+y+z pass, z-slab)", "achieved": ALG_BYTES_MATMULT * value / world,
                "peak": peak, "unit": "GB/s", "frac": ALG_BYTES_MATMULT * value / world / peak, "traffic": None,
                "peak_source": peak_src}

    # end to end through the host-pointer C-ABI call (what the Fortran shim calls), pinned buffers
    e2e = None
    if not args.no_e2e:
        import ctypes

        fh = torch.empty((nzl, n, n), dtype=torch.float64).pin_memory()
        oh = torch.empty((nzl, n, n), dtype=torch.float64).pin_memory()
        fh.copy_(f)
        torch.cuda.synchronize()
        ksteps = max(3, min(args.steps, 10))
    if not args.no_e2e and world == 1:
        d3 = _lib._d3(*dx)
        pf = ctypes.cast(fh.data_ptr(), _lib._dp)
        po = ctypes.cast(oh.data_ptr(), _lib._dp)
        for _ in range(2):
            pbx.check(pbx.LIB.pbx_lapl_host(n, n, nzl, pf, d3, po, pbx.MODE_FAST))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            pbx.check(pbx.LIB.pbx_lapl_host(n, n, nzl, pf, d3, po, pbx.MODE_FAST))
        dt = (time.perf_counter() - t0) / ksteps
        assert torch.equal(oh.to(dev), out), "host-pointer path disagrees with the device path"
        e2e = {"value": ndof_total / dt / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": int(8 * ndof_total),
               "d2h_bytes_per_step": int(8 * ndof_total), "ms_per_step": dt * 1e3, "steps": ksteps,
               "api": "pbx_lapl_host (pinned host buffers)",
               "floor": "one apply needs the whole field before the first output plane exists (z lines), so copy-in and "
                        "copy-out of ONE field cannot overlap: 2 x 1 GiB over the host link + 2 ms of compute"}
        # several fields per call, double-buffered (copy-in / compute / copy-out of consecutive fields overlap:
        # the host link runs in both directions at once)
        try:
            oh2 = torch.empty((nzl, n, n), dtype=torch.float64).pin_memory()
            pin = (ctypes.c_void_p * ksteps)(*([fh.data_ptr()] * ksteps))
            pout = (ctypes.c_void_p * ksteps)(*[(oh if k % 2 == 0 else oh2).data_ptr() for k in range(ksteps)])
            pbx.check(pbx.LIB.pbx_lapl_host_batch(n, n, nzl, 2, pin, d3, pout, pbx.MODE_FAST))
            t0 = time.perf_counter()
            pbx.check(pbx.LIB.pbx_lapl_host_batch(n, n, nzl, ksteps, pin, d3, pout, pbx.MODE_FAST))
            dtb = (time.perf_counter() - t0) / ksteps
            same = torch.equal(oh.to(dev), out) and torch.equal(oh2.to(dev), out)
            e2e["batch"] = {"value": ndof_total / dtb / 1e9, "ms_per_field": dtb * 1e3, "fields": ksteps,
                            "api": "pbx_lapl_host_batch (double-buffered, pinned host buffers)",
                            "matches_device_path": bool(same)}
            del oh2
        except Exception as exc:
            e2e["batch"] = {"error": repr(exc)}
        pbx.LIB.pbx_host_cache_clear()
    elif not args.no_e2e:
        # N > 1: every rank stages ITS slab from pinned host memory, applies the distributed operator and reads
        # its slab of the result back (the host-pointer C calls carry no communicator; this is the same
        # sequence through the handle API).  Wall clock between barriers, max over ranks.
        def one():
            f.copy_(fh, non_blocking=True)
            h.lapl(f, out)
            oh.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            one()
        dt = (time.perf_counter() - t0) / ksteps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
        e2e = {"value": ndof_total / dt / 1e9, "unit": "GDoF/s", "h2d_bytes_per_step": int(8 * ndof_total),
               "d2h_bytes_per_step": int(8 * ndof_total), "ms_per_step": dt * 1e3, "steps": ksteps,
               "api": "per rank: pinned slab -> device, pbx_lapl_device over the communicator, slab -> pinned host"}
    if not args.no_e2e:
        del fh, oh

    # CG time-to-rtol on a manufactured smooth solution (S4), device resident
    cg = None
    if not args.no_cg:
        hh = 2 * np.pi / n
        c = (torch.arange(n, dtype=torch.float64, device=dev) + 0.5) * hh
        cz = c[rank * nzl:(rank + 1) * nzl]
        u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(cz)[:, None, None]).contiguous()
        h2 = pbx.Handle(n, n, nzl, (hh,) * 3, device=local, comm=comm)
        h2.use_current_stream()
        b = h2.lapl(u)
        x = h2.empty()
        del u
        # warm-up solve (5 iterations): workspace allocation and, for N>1, NCCL's lazy
        # peer-to-peer connection set-up are not part of the time-to-solution
        h2.cg_solve(b, x, rtol=args.cg_rtol, maxit=5)
        barrier()
        l1 = h2.launches
        t0 = time.perf_counter()
        x, its, rnorm, reason, hist = h2.cg_solve(b, x, rtol=args.cg_rtol, maxit=args.cg_maxit)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        # independent check of the answer: true residual ||A x - b|| / ||b|| (outside the timed region)
        rres = h2.lapl(x) - b
        num, den = (rres * rres).sum(), (b * b).sum()
        if world > 1:
            dist.all_reduce(num)
            dist.all_reduce(den)
        true_rel = float(torch.sqrt(num / den).item())
        del rres
        cg = {"rhs": "S4 manufactured: b = A exp(sin x + sin y + sin z), L = 2 pi", "rtol": args.cg_rtol,
              "true_residual_rel": true_rel,
              "time_s": dt, "its": its, "reason": reason, "rnorm_rel": rnorm / hist[0] if hist[0] else 0.0,
              "ms_per_it": dt / max(1, its) * 1e3, "GDoF_it_per_s": ndof_total * its / dt / 1e9,
              # 152 B/DoF per iteration (80 MatMult + 72 vectors, p.w counted as free).  The build
              # moves 88 (the z pass reads p for the fused dot) + 64 (x += a p rides with the p update) = 152.
              "alg_bytes_per_dof_it": 152.0, "frac_of_hbm_peak": 152.0 * ndof_total * its / dt / 1e9 / peak / world,
              "gpu_launches": h2.launches - l1}
        # the same solve with the multigrid preconditioner on the 2nd-order star (the
        # role of the reference's `-pc_type gamg` on P).  Reported beside the headline, which stays the
        # reference's unpreconditioned configuration.  On slabs (N > 1: distributed levels with one-plane
        # halo exchanges, coarse levels all-gathered) the leg is opt-in, PBX_BENCH_MG_SLABS=1, until it
        # has run on the GPUs once: it was written after the round's GPU budget was spent.
        if world == 1 or os.environ.get("PBX_BENCH_MG_SLABS") == "1":
            try:
                from poissbox_b200 import _lib as _pl

                h2.set_pc(_pl.PC_MG, 2)
                h2.cg_solve(b, x, rtol=args.cg_rtol, maxit=2)
                torch.cuda.synchronize()
                l1 = h2.launches
                t0 = time.perf_counter()
                x, its, rnorm, reason, hist = h2.cg_solve(b, x, rtol=args.cg_rtol, maxit=args.cg_maxit)
                torch.cuda.synchronize()
                dtp = time.perf_counter() - t0
                rres = h2.lapl(x) - b
                num, den = (rres * rres).sum(), (b * b).sum()
                if world > 1:
                    t = torch.tensor([dtp], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dtp = t.item()
                    dist.all_reduce(num)
                    dist.all_reduce(den)
                true_rel = float(torch.sqrt(num / den).item())
                del rres
                cg["multigrid_pc"] = {"pc": "V(2,2) geometric multigrid on the 2nd-order star, damped Jacobi",
                                      "stopping": "PETSc's default: preconditioned norm ||M^-1 r|| <= rtol ||M^-1 b||",
                                      "time_s": dtp, "its": its, "reason": reason,
                                      "rnorm_rel": rnorm / hist[0] if hist[0] else 0.0,
                                      "true_residual_rel": true_rel, "ms_per_it": dtp / max(1, its) * 1e3,
                                      "gpu_launches": h2.launches - l1}
                # the like-for-like figure: the same solve carried on until the TRUE residual ||A x - b|| / ||b|| is
                # below the tolerance the unpreconditioned solve reaches (M ~ P^-1 weights smooth error, so the
                # preconditioned norm is 2-3 orders of magnitude ahead of the true residual)
                for rt in (1e-10, 3e-11, 1e-11, 3e-12):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    x, its, rnorm, reason, hist = h2.cg_solve(b, x, rtol=rt, maxit=args.cg_maxit)
                    torch.cuda.synchronize()
                    dte = time.perf_counter() - t0
                    rres = h2.lapl(x) - b
                    num = (rres * rres).sum()
                    if world > 1:
                        t = torch.tensor([dte], dtype=torch.float64, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        dte = t.item()
                        dist.all_reduce(num)
                    tr = float(torch.sqrt(num / den).item())
                    del rres
                    if tr <= args.cg_rtol or rt == 3e-12:
                        cg["multigrid_pc"]["equal_true_residual"] = {
                            "rtol_on_preconditioned_norm": rt, "its": its, "reason": reason, "time_s": dte,
                            "true_residual_rel": tr, "speedup_vs_pc_none": dt / dte}
                        break
            except Exception as exc:   # the optional leg must never cost the headline line
                cg["multigrid_pc"] = {"error": repr(exc)}
        h2.close()
        del b, x

    # BASELINE configs[2] (256^3 on one GPU) and the full-spectrum problem at scale; end-to-end time-to-solution
    # through the host-pointer call (b in once, x out once)
    if cg is not None and world == 1 and not args.no_cg and not args.quick:
        cg["other_configs"] = []
        for nn, kind in ((256, "S4"), (256, "S3"), (128, "S3")):
            try:
                cg["other_configs"].append(cg_case(pbx, torch, nn, kind, args.cg_rtol, local))
            except Exception as exc:
                cg["other_configs"].append({"grid": [nn] * 3, "rhs": kind, "error": repr(exc)})
        if e2e is not None:
            try:
                import ctypes

                hh = 2 * np.pi / n
                c = (torch.arange(n, dtype=torch.float64, device=dev) + 0.5) * hh
                u = torch.exp(torch.sin(c)[None, None, :] + torch.sin(c)[None, :, None] + torch.sin(c)[:, None, None]).contiguous()
                hb = pbx.Handle(n, n, n, (hh,) * 3, device=local)
                bh = hb.lapl(u).cpu().pin_memory()
                hb.close()
                del u, hb
                xh = torch.empty_like(bh).pin_memory()
                its_, why_, rn_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
                t0 = time.perf_counter()
                pbx.check(pbx.LIB.pbx_cg_solve_host(n, n, n, _lib._d3(hh, hh, hh), ctypes.cast(bh.data_ptr(), _lib._dp),
                                                    ctypes.cast(xh.data_ptr(), _lib._dp), args.cg_rtol, 1e-50, args.cg_maxit,
                                                    pbx.MODE_FAST, ctypes.byref(its_), ctypes.byref(rn_), ctypes.byref(why_),
                                                    None, 0))
                dth = time.perf_counter() - t0
                e2e["cg"] = {"api": "pbx_cg_solve_host (b in once, x out once, pinned host buffers; includes workspace allocation)",
                             "time_s": dth, "its": its_.value, "reason": why_.value,
                             "h2d_bytes": int(8 * ndof_total), "d2h_bytes": int(8 * ndof_total),
                             "device_resident_time_s": cg["time_s"]}
                pbx.LIB.pbx_host_cache_clear()
                del bh, xh
            except Exception as exc:
                e2e["cg"] = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_block()
        if not args.quick:
            try:
                cpu["cg_iteration_parity"] = cg_vs_oracle(pbx, 128, 1e-5)
            except Exception as exc:
                cpu["cg_iteration_parity"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GDoF/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"compact Laplacian apply (FAST schedule), {n}^3 fp64 periodic box, S2 random field "
                                   f"U[-1,1], dx = 1/{n}; N>1: z-slabs of {nzl} planes",
                       "grid": [n, n, n], "local_brick": [n, n, nzl], "l2": "field (1 GiB at 512^3) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"zslab{world}",
                       "rank_sync": ("n/a" if world == 1 else
                                     "peer boards (flag barrier + all-reduce in the CG's reduction kernel, no NCCL per iteration)"
                                     if pbx.LIB.pbx_peer_sync_active(h._h) == 1 else
                                     "peer stores of the boundary messages + ncclAllReduce (barrier, CG scalars)")},
            "parity": parity, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "cg": cg,
        }
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        if comm:
            pbx.LIB.pbx_comm_destroy(comm)
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE configs[4] on one GPU: tools/sweep_bench.py's measurements as ONE JSON line (the driver-visible form)"""
    import io
    from contextlib import redirect_stdout

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_bench

    buf = io.StringIO()
    with redirect_stdout(buf):
        recs = sweep_bench.run(logn=27 if not args.quick else 24)
    peak, src = hbm_peak()
    best = {}
    for r in recs:
        key = r["op"] + ("/" + r["layout"] if "layout" in r else "")
        f = r.get("frac_hbm")
        if f is not None:
            lo, hi = best.get(key, (1e9, 0.0))
            best[key] = (min(lo, f), max(hi, f))
    print(json.dumps({"metric": "config5: batched tridsol + compact grad/div/Laplacian sweeps, lines 64-2048, one B200",
                      "unit": "fraction of the measured HBM peak on the ALGORITHMIC bytes of SURVEY 8(d) "
                              "(lapl 80 B/DoF, grad/div 32 B/DoF, tdma 48 B/point)",
                      "peak_GBs": peak, "peak_source": src, "n_gpus": 1, "dtype": "f64", "data": "synthetic",
                      "frac_range": {k: [round(v[0], 4), round(v[1], 4)] for k, v in best.items()},
                      "results": recs}), flush=True)


def main():
    args = parse()
    if args.workload == "config5":
        run_config5(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
