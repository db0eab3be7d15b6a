import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


@pytest.fixture(scope="session")
def oracle_built():
    """oracle/libminife_oracle.so (gcc, plain C)."""
    so = os.path.join(ROOT, "oracle", "libminife_oracle.so")
    src = [os.path.join(ROOT, "oracle", f) for f in ("minife_oracle.c", "minife_oracle.h")]
    if not _newer(so, src):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return so


@pytest.fixture(scope="session")
def harness():
    """tests/libcpu_harness.so: the kernels' host/device-shared logic compiled for the CPU (test-only)."""
    import ctypes as C
    so = os.path.join(ROOT, "tests", "libcpu_harness.so")
    csrc = os.path.join(ROOT, "minife_b200", "csrc")
    src = [os.path.join(ROOT, "tests", "cpu_harness.cpp")] + [os.path.join(csrc, f) for f in
                                                               ("mesh.h", "hex8.h", "assembly.h", "exact_dot.h", "halo_maps.h")]
    if not _newer(so, src):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-w", "-shared", "-fPIC", "-o", so, src[0]], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="session")
def product_lib():
    """minife_b200/libminife_b200.so, built if missing (nvcc cross-compiles without a GPU)."""
    import minife_b200
    if not os.path.exists(minife_b200.library_path()):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "minife_b200", "csrc")], check=True)
    return minife_b200.lib()
