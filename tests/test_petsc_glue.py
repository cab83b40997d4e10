"""petsc/pbx_matshell.c (the MATSHELL MatMult over VECCUDA arrays, the counterpart of mfmult,
src/poissbox.f90:300-322) compiled against a mock of the PETSc calls it makes
(tests/petsc_mock/petsc_mock.h; this image has no PETSc): it must compile warning-free everywhere,
refuse to run without a GPU, and on a GPU its MatMult must reproduce pbx_lapl_host bit for bit and
its CG must solve A x = b."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK = os.path.join(HERE, "petsc_mock")
EXE = os.path.join(MOCK, "_build", "test_matshell")


def build():
    subprocess.check_call(["make", "-s", "-C", MOCK])


def test_petsc_glue_compiles_and_fails_loudly_without_gpu():
    build()
    import poissbox_b200 as pbx

    if pbx.LIB.pbx_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr


def test_petsc_glue_on_cpu_harness():
    """the same driver, compiled as C++ against the CPU kernel-logic harness (tests/emu): the
    glue's MatMult and CG paths execute end to end without a GPU"""
    import emu_lib

    emu_lib.load()
    emu = os.path.join(HERE, "emu")
    exe = os.path.join(MOCK, "_build", "test_matshell_emu")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-w", "-I" + os.path.join(emu, "include"),
                           "-include", "pbx_emu.h", "-I" + MOCK, "-I" + os.path.join(HERE, "..", "include"),
                           os.path.join(MOCK, "test_matshell.c"), "-o", exe,
                           "-L" + os.path.join(emu, "_build"), "-lpbx_emu",
                           "-Wl,-rpath," + os.path.join(emu, "_build")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_petsc_glue_matmult_and_cg():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
