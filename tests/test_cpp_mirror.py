"""The C++ host-side mirror (include/pbx_host.hpp): builds everywhere; on a GPU the C++
restatements of the reference's test programs must pass, and a length mismatch must end with the
reference's `stop 7`."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")


def build():
    subprocess.check_call(["make", "-s", "-C", CPP])


def test_cpp_mirror_builds_and_fails_loudly_without_gpu():
    build()
    import poissbox_b200 as pbx

    if pbx.LIB.pbx_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu tests")
    r = subprocess.run([os.path.join(CPP, "_build", "test_lapl")], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["test_lapl"], ["test_lapl", "ref"], ["test_tdma_periodic"]])
def test_cpp_reference_tests(args):
    build()
    r = subprocess.run([os.path.join(CPP, "_build", args[0])] + args[1:], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout and "failed" not in r.stdout


def test_cpp_size_mismatch_is_stop_7():
    build()
    r = subprocess.run([os.path.join(CPP, "_build", "test_tdma_periodic"), "mismatch"], capture_output=True, text=True)
    assert r.returncode == 7
    assert "periodic gradient is same length as field" in r.stdout
