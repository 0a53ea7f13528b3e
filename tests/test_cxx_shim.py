"""The reference-side binding as a C++ translation unit (tests/cxx): compiled with g++ -std=c++20 against
include/floxer_gpu.h and linked with libfloxer_gpu.so.  On the CPU it must compile, link and resolve every call; on the
GPU it runs the reference's own known-answer cases (test/alignment_test.cpp:7-30, test/verification_test.cpp:11-123)
through definitions of alignment::align and verification::query_verifier::verify with the reference's signatures."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BINARY = os.path.join(HERE, "cxx", "shim_test.bin")


def build_binary() -> str:
    from floxer_b200 import build
    build.build_native()
    lib_dir = os.path.join(ROOT, "floxer_b200")
    srcs = [os.path.join(HERE, "cxx", f) for f in ("shim_test.cpp", "floxer_shim.hpp", "reference_stubs.hpp")] + [os.path.join(ROOT, "include", "floxer_gpu.h")]
    if not os.path.exists(BINARY) or any(os.path.getmtime(s) > os.path.getmtime(BINARY) for s in srcs):
        cmd = ["g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), srcs[0], "-o", BINARY,
               "-L", lib_dir, "-lfloxer_gpu", f"-Wl,-rpath,{lib_dir}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return BINARY


def test_shim_compiles_and_links():
    r = subprocess.run([build_binary(), "--link-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "floxer_b200" in r.stdout


@pytest.mark.gpu
def test_reference_cases_through_the_shim():
    r = subprocess.run([build_binary()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
