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


@pytest.mark.gpu
def test_petsc_glue_matmult_and_cg():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
