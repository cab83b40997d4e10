"""CPU: schedule independence of the kernels on the kernel-logic harness (tests/emu).  A fiber runs
undisturbed between two switch points (barriers, shuffles, mbarrier waits), so the order in which
the harness runs the threads of a CTA decides which of two unsynchronised shared-memory accesses
between the same barriers comes first.  Kernels free of such races give the same bits whatever
the order: ascending, descending and two pseudo-random orders (reshuffled every round) must agree
for the Laplacian (TMA and generic kernels, long lines), grad / div / interp, the REFERENCE
schedule, the slab decomposition and both CG loops."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def run(sched):
    import emu_lib

    emu_lib.load()   # build once, in this process
    env = dict(os.environ, PBX_EMU_SCHED=str(sched))
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_sched_worker.py")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout.strip().splitlines()[-1]


def test_emu_results_do_not_depend_on_thread_order():
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(4) as ex:
        digests = list(ex.map(run, (0, 1, 2, 7)))
    assert len(set(digests)) == 1, digests
