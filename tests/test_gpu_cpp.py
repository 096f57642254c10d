"""The C++ host-side mirror (include/quaru_b200.hpp): compiled with g++ against the C-ABI library.
The host-only part (constructors, circuit builder) runs on CPU; the register part needs a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(engine):
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "test_register")
    lib = os.path.join(ROOT, "quaru_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_register.cpp"), "-o", exe, "-L", lib, "-lquaru_b200",
                           f"-Wl,-rpath,{lib}"])
    return exe


def test_cpp_mirror_host_only(engine):
    out = subprocess.run([_build(engine), "--host-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "HOST_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(engine):
    out = subprocess.run([_build(engine)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "CPP_OK" in out.stdout, out.stdout + out.stderr
